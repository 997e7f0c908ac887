"""ORACLE (test infrastructure only) for SURVEY section 8f rank 4, BASELINE configs[4]: the Sim(3) pose-graph optimisation of
Map::runPoseGraphOptimization (src/Map.cc:75-165), restated in numpy / scipy from the cited lines:

  * vertices VertexSim3 (include/internal/PR.h:57-76): estimate Sim3 (Scw), oplus: estimate = Sim3::exp(update) * estimate;
  * edges EdgeSim3 (PR.h:151-179): error = log(measurement^-1 * v1 * v2^-1) in R^7, information 7x7;
  * Jacobians: g2o's numeric BaseBinaryEdge::linearizeOplus (thirdparty/g2o/g2o/core/base_binary_edge.hpp:131-148): central
    differences with delta = 1e-9 through oplus, one tangent dimension at a time;
  * quadratic form (base_binary_edge.hpp:59-106): H_ii += Ji^T O Ji, H_ij += Ji^T O Jj, H_jj += Jj^T O Jj, b_i += -Ji^T O e, ...;
  * OptimizationAlgorithmGaussNewton: solve H dx = b (LinearSolverEigen: sparse LDLT -- here scipy's sparse LU, the same exact
    solve), oplus every non-fixed vertex, 25 iterations (Map.cc:141), the current keyframe's vertex fixed (Map.cc:109-111);
  * Sim3 / RxSO3 / SO3 exp and log as in thirdparty/sophus/sim3.hpp:418-427,572-589,609-646, rxso3.hpp:416-425,553-562,
    so3.hpp:343-369,491-531 (quaternion with norm = scale; tangent = [upsilon, omega, sigma]).

g2o and Eigen are not on this machine and the reference ships no fixture for this path: parity UNPINNED -- this file is the checker,
itself checked against invariants (exp/log round trips, zero error at the ground truth, convergence to the noise-free graph)."""
import numpy as np

EPS = 1e-10          # SophusConstants<double>::epsilon()
DELTA = 1e-9         # g2o numeric Jacobian step


# ---- quaternions (w, x, y, z), arrays [..., 4]; a Sim3 is (q with |q| = scale, t)
def qmul(a, b):
    aw, ax, ay, az = np.moveaxis(a, -1, 0); bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def qrot(q, v):
    """(scaled) rotation of v by the non-unit quaternion q: s R v with s = |q| (RxSO3::operator*)."""
    s = np.linalg.norm(q, axis=-1, keepdims=True)
    u = q / s
    w = u[..., :1]; r = u[..., 1:]
    t = 2.0 * np.cross(r, v)
    return s * (v + w * t + np.cross(r, t))


def sim3_mul(a, b):
    return qmul(a[0], b[0]), a[1] + qrot(a[0], b[1])


def sim3_inv(a):
    q, t = a
    qi = qconj(q) / np.sum(q * q, axis=-1, keepdims=True)
    return qi, -qrot(qi, t)


def hat(w):
    z = np.zeros_like(w[..., 0])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1), np.stack([w[..., 2], z, -w[..., 0]], -1), np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def calc_w(theta, sigma, scale, omega):
    """sim3.hpp:609-646"""
    Om = hat(omega); Om2 = Om @ Om
    small_s = np.abs(sigma) < EPS; small_t = np.abs(theta) < EPS
    th2 = theta * theta
    with np.errstate(divide="ignore", invalid="ignore"):
        C = np.where(small_s, 1.0, (scale - 1.0) / np.where(small_s, 1.0, sigma))
        A_ss = np.where(small_t, 0.5, (1.0 - np.cos(theta)) / np.where(small_t, 1.0, th2))
        B_ss = np.where(small_t, 1.0 / 6.0, (theta - np.sin(theta)) / np.where(small_t, 1.0, th2 * theta))
        s2 = sigma * sigma
        A_st = ((sigma - 1.0) * scale + 1.0) / np.where(small_s, 1.0, s2)
        B_st = ((0.5 * s2 - sigma + 1.0) * scale) / np.where(small_s, 1.0, s2 * sigma)
        a = scale * np.sin(theta); b = scale * np.cos(theta); c = th2 + s2
        A_g = (a * sigma + (1.0 - b) * theta) / np.where(small_t | small_s, 1.0, theta * c)
        B_g = (C - ((b - 1.0) * sigma + a * theta) / np.where(c == 0, 1.0, c)) * 1.0 / np.where(small_t, 1.0, th2)
    A = np.where(small_s, A_ss, np.where(small_t, A_st, A_g))
    B = np.where(small_s, B_ss, np.where(small_t, B_st, B_g))
    I = np.eye(3)
    return A[..., None, None] * Om + B[..., None, None] * Om2 + C[..., None, None] * I


def sim3_exp(a):
    ups, om, sigma = a[..., 0:3], a[..., 3:6], a[..., 6]
    th2 = np.sum(om * om, -1); theta = np.sqrt(th2); half = 0.5 * theta
    small = theta < EPS
    with np.errstate(divide="ignore", invalid="ignore"):
        imag = np.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, np.sin(half) / np.where(small, 1.0, theta))
    real = np.where(small, 1.0 - 0.5 * th2 + th2 * th2 / 384.0, np.cos(half))
    scale = np.exp(sigma)
    q = np.concatenate([real[..., None], imag[..., None] * om], -1) * scale[..., None]
    W = calc_w(theta, sigma, scale, om)
    return q, np.einsum("...ij,...j->...i", W, ups)


def sim3_log(T):
    q, t = T
    scale = np.linalg.norm(q, axis=-1)
    sigma = np.log(scale)
    u = q / scale[..., None]
    n2 = np.sum(u[..., 1:] ** 2, -1); n = np.sqrt(n2); w = u[..., 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        f_small = 2.0 / w - 2.0 * n2 / (w * w * w)
        f_w0 = np.where(w > 0, np.pi, -np.pi) / np.where(n == 0, 1.0, n)
        f_gen = 2.0 * np.arctan(n / np.where(w == 0, 1.0, w)) / np.where(n == 0, 1.0, n)
    f = np.where(n < EPS, f_small, np.where(np.abs(w) < EPS, f_w0, f_gen))
    theta = f * n
    omega = f[..., None] * u[..., 1:]
    W = calc_w(theta, sigma, scale, omega)
    ups = np.linalg.solve(W, t[..., None])[..., 0]
    return np.concatenate([ups, omega, sigma[..., None]], -1)


def edge_error(meas_inv, Vi, Vj):
    """EdgeSim3::computeError (PR.h:162-166)"""
    return sim3_log(sim3_mul(sim3_mul(meas_inv, Vi), sim3_inv(Vj)))


def linearize(q, t, ei, ej, mq, mt, info):
    """errors e [nE,7], numeric Jacobians Ji, Jj [nE,7,7] (column d = d error / d update_d), chi2."""
    Vi, Vj = (q[ei], t[ei]), (q[ej], t[ej])
    Minv = sim3_inv((mq, mt))
    e = edge_error(Minv, Vi, Vj)
    nE = len(ei)
    Ji = np.zeros((nE, 7, 7)); Jj = np.zeros((nE, 7, 7))
    for d in range(7):
        up = np.zeros(7); up[d] = DELTA
        Ep, Em = sim3_exp(up), sim3_exp(-up)
        bc = lambda T: (np.broadcast_to(T[0], (nE, 4)), np.broadcast_to(T[1], (nE, 3)))
        Ji[:, :, d] = (edge_error(Minv, sim3_mul(bc(Ep), Vi), Vj) - edge_error(Minv, sim3_mul(bc(Em), Vi), Vj)) / (2 * DELTA)
        Jj[:, :, d] = (edge_error(Minv, Vi, sim3_mul(bc(Ep), Vj)) - edge_error(Minv, Vi, sim3_mul(bc(Em), Vj))) / (2 * DELTA)
    chi2 = float(np.einsum("ei,eij,ej->", e, info, e))
    return e, Ji, Jj, chi2


def optimize(q, t, ei, ej, mq, mt, info, fixed, iterations=25):
    """Gauss-Newton as g2o runs it. Returns (q, t, chi2 before every iteration + after the last)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    q, t = q.copy(), t.copy()
    nV = len(q)
    free = np.ones(nV, bool); free[fixed] = False
    idx = -np.ones(nV, np.int64); idx[free] = np.arange(free.sum())
    chis = []
    for it in range(iterations):
        e, Ji, Jj, chi2 = linearize(q, t, ei, ej, mq, mt, info)
        chis.append(chi2)
        OJi = info @ Ji; OJj = info @ Jj
        Hii = np.einsum("eri,erj->eij", Ji, OJi); Hij = np.einsum("eri,erj->eij", Ji, OJj); Hjj = np.einsum("eri,erj->eij", Jj, OJj)
        bi = -np.einsum("eri,er->ei", OJi, e); bj = -np.einsum("eri,er->ei", OJj, e)
        n = int(free.sum()) * 7
        rows, cols, vals = [], [], []
        b = np.zeros(n)
        r7 = np.arange(7)
        def add(bi_, bj_, blocks, m):
            base_r = (idx[bi_][m] * 7)[:, None, None] + r7[None, :, None]
            base_c = (idx[bj_][m] * 7)[:, None, None] + r7[None, None, :]
            rows.append(np.broadcast_to(base_r, blocks[m].shape).ravel()); cols.append(np.broadcast_to(base_c, blocks[m].shape).ravel()); vals.append(blocks[m].ravel())
        fi, fj = free[ei], free[ej]
        add(ei, ei, Hii, fi); add(ej, ej, Hjj, fj); add(ei, ej, Hij, fi & fj); add(ej, ei, np.transpose(Hij, (0, 2, 1)), fi & fj)
        np.add.at(b, (idx[ei][fi] * 7)[:, None] + r7[None, :], bi[fi]); np.add.at(b, (idx[ej][fj] * 7)[:, None] + r7[None, :], bj[fj])
        H = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
        dx = spl.spsolve(H, b).reshape(-1, 7)
        E = sim3_exp(dx)
        nq, nt = sim3_mul(E, (q[free], t[free]))
        q[free], t[free] = nq, nt
    chis.append(linearize(q, t, ei, ej, mq, mt, info)[3])
    return q, t, np.array(chis)


def make_graph(n_kf=5000, n_loop=10000, seed=0, noise=0.02, drift=1e-4):
    """BASELINE configs[4]: a KF chain on a circle with odometry edges to the previous 2 keyframes, n_loop loop edges
    between keyframes >= 100 apart; measurements = ground-truth relative Sim3 (Tcr = S_i * S_j^-1); initial estimates = ground truth
    with accumulated scale drift (drift per KF) and pose noise; information = I7."""
    rng = np.random.default_rng(seed)
    ang = np.linspace(0, 4 * np.pi, n_kf)
    rad = 10.0
    # camera-to-world positions on a circle, looking along the tangent; Scw = inverse
    twc = np.stack([rad * np.cos(ang), rad * np.sin(ang), 0.3 * np.sin(3 * ang)], -1)
    yaw = ang + np.pi / 2
    qwc = np.stack([np.cos(yaw / 2), np.zeros(n_kf), np.zeros(n_kf), np.sin(yaw / 2)], -1)
    gq, gt = sim3_inv((qwc, twc))           # ground truth Scw, scale 1
    ei, ej = [], []
    for k in range(1, n_kf):
        for d in (1, 2):
            if k - d >= 0:
                ei.append(k); ej.append(k - d)
    li = rng.integers(0, n_kf, 4 * n_loop); lj = rng.integers(0, n_kf, 4 * n_loop)
    m = np.abs(li - lj) >= min(100, n_kf // 4)
    li, lj = li[m][:n_loop], lj[m][:n_loop]
    ei = np.concatenate([np.array(ei), li]); ej = np.concatenate([np.array(ej), lj])
    mq, mt = sim3_mul((gq[ei], gt[ei]), sim3_inv((gq[ej], gt[ej])))      # Tcr: error log(M^-1 Vi Vj^-1) = 0 at the ground truth
    # initial estimates: scale drift along the chain + small pose noise (left-multiplied tangent perturbation)
    sig = np.cumsum(np.full(n_kf, drift)) - drift
    pert = np.concatenate([rng.normal(0, noise, (n_kf, 3)), rng.normal(0, noise * 0.2, (n_kf, 3)), sig[:, None]], -1)
    pert[-1] = 0.0; pert[-1, 6] = 0.0
    q0, t0 = sim3_mul(sim3_exp(pert), (gq, gt))
    info = np.broadcast_to(np.eye(7), (len(ei), 7, 7)).copy()
    return dict(q=q0, t=t0, ei=ei.astype(np.int32), ej=ej.astype(np.int32), mq=mq, mt=mt, info=info, fixed=n_kf - 1, gq=gq, gt=gt)
