"""Seeded synthetic keyframe windows shared by the tests, the bench and the oracle driver.

Nothing here touches the GPU or the oracle: it only produces the *inputs* (images, poses, points,
residual lists) as numpy arrays, in the layouts the C ABI (include/ldso_b200.h) consumes.

Scene: a procedural texture T(X,Y) = 128 + sum_k A_k sin(f_k (X cos th_k + Y sin th_k) + ph_k) painted on a
near half-plane Z = 2 m (world X <= 0) in front of a full far plane Z = 5 m; images are rendered exactly by
ray/plane intersection per pixel, then I_i = exp(a_i) T + b_i (exposure tau = 1) clipped to [0, 255].
Pyramids and gradients follow the reference's FrameHessian::makeImages (src/internal/FrameHessian.cc:44-98);
point colours/weights follow ImmaturePoint's constructor (src/internal/ImmaturePoint.cc:14-39).
"""
from __future__ import annotations

import dataclasses
import numpy as np

SCALE_XI_ROT = 1.0
SCALE_XI_TRANS = 0.5
SCALE_A = 10.0
SCALE_B = 1000.0
PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)
OUTLIER_TH_SUM_COMPONENT = 50.0 * 50.0


def pyr_levels_for(w: int, h: int, max_levels: int = 6) -> int:
    """setGlobalCalib's level count (src/internal/GlobalCalib.cc:20-28)."""
    wl, hl, n = w, h, 1
    while wl % 2 == 0 and hl % 2 == 0 and wl * hl > 5000 and n < max_levels:
        wl //= 2
        hl //= 2
        n += 1
    return n


def so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


class Texture:
    def __init__(self, seed=1234, n=48):
        rng = np.random.default_rng(seed)
        self.A = rng.uniform(1.0, 6.0, n)
        self.f = np.exp(rng.uniform(np.log(2.0), np.log(60.0), n))
        self.th = rng.uniform(0, 2 * np.pi, n)
        self.ph = rng.uniform(0, 2 * np.pi, n)

    def __call__(self, X, Y):
        out = np.full(X.shape, 128.0, dtype=np.float64)
        c, s = np.cos(self.th), np.sin(self.th)
        for k in range(len(self.A)):
            out += self.A[k] * np.sin(self.f[k] * (X * c[k] + Y * s[k]) + self.ph[k])
        return out


def scene_depth(Rcw, tcw, K, px, py):
    """Ray/plane intersection for pixels (px,py) of a camera with worldToCam (Rcw,tcw).
    Returns world hit (X,Y), camera-frame depth Z_c and the far-plane flag."""
    fx, fy, cx, cy = K
    Rwc = Rcw.T
    c = -Rwc @ tcw
    d_cam = np.stack([(px - cx) / fx, (py - cy) / fy, np.ones_like(px)], axis=-1)
    d_w = d_cam @ Rwc.T
    s_near = (2.0 - c[2]) / d_w[..., 2]
    Xn = c[0] + s_near * d_w[..., 0]
    far = Xn > 0.0
    s = np.where(far, (5.0 - c[2]) / d_w[..., 2], s_near)
    X = c[0] + s * d_w[..., 0]
    Y = c[1] + s * d_w[..., 1]
    return X, Y, s, far  # d_cam has z = 1, so the camera-frame depth equals s


def render(tex, Rcw, tcw, K, w, h, a=0.0, b=0.0):
    px, py = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    X, Y, _, far = scene_depth(Rcw, tcw, K, px, py)
    T = tex(np.where(far, X + 7.3, X), np.where(far, Y - 1.1, Y))
    img = np.exp(a) * T + b
    return np.clip(img, 0.0, 255.0).astype(np.float32)


def make_images(color: np.ndarray, levels: int):
    """FrameHessian::makeImages (FrameHessian.cc:44-98): list of (h_l, w_l, 3) float32 (I, dx, dy)."""
    h, w = color.shape
    out = []
    I = color.astype(np.float32)
    for lvl in range(levels):
        wl, hl = w >> lvl, h >> lvl
        if lvl > 0:
            P = out[lvl - 1][:, :, 0]
            I = np.float32(0.25) * (((P[0:2 * hl:2, 0:2 * wl:2] + P[0:2 * hl:2, 1:2 * wl:2]) + P[1:2 * hl:2, 0:2 * wl:2])
                                    + P[1:2 * hl:2, 1:2 * wl:2])
            I = I.astype(np.float32)
        flat = np.ascontiguousarray(I).reshape(-1)
        dI = np.zeros((hl * wl, 3), dtype=np.float32)
        dI[:, 0] = flat
        lo, hi = wl, wl * (hl - 1)
        dx = np.float32(0.5) * (flat[lo + 1:hi + 1] - flat[lo - 1:hi - 1])
        dy = np.float32(0.5) * (flat[lo + wl:hi + wl] - flat[lo - wl:hi - wl])
        dx = np.where(np.isnan(dx) | (np.abs(dx) > 255.0), np.float32(0), dx)
        dy = np.where(np.isnan(dy) | (np.abs(dy) > 255.0), np.float32(0), dy)
        dI[lo:hi, 1] = dx
        dI[lo:hi, 2] = dy
        out.append(dI.reshape(hl, wl, 3))
    return out


def sample_bilin(dI0: np.ndarray, x: np.ndarray, y: np.ndarray):
    """getInterpolatedElement33BiLin (GlobalFuncs.h:185-207) on channel 0, vectorised, float32."""
    x = x.astype(np.float32)
    y = y.astype(np.float32)
    ix = x.astype(np.int32)
    iy = y.astype(np.int32)
    I = dI0[:, :, 0]
    tl, tr = I[iy, ix], I[iy, ix + 1]
    bl, br = I[iy + 1, ix], I[iy + 1, ix + 1]
    dx = x - ix.astype(np.float32)
    dy = y - iy.astype(np.float32)
    one = np.float32(1)
    topInt = dx * tr + (one - dx) * tl
    botInt = dx * br + (one - dx) * bl
    leftInt = dy * bl + (one - dy) * tl
    rightInt = dy * br + (one - dy) * tr
    return dx * rightInt + (one - dx) * leftInt, rightInt - leftInt, botInt - topInt


@dataclasses.dataclass
class Window:
    w: int
    h: int
    levels: int
    K: np.ndarray                 # (4,) f64 fx fy cx cy  (== CalibHessian::value_scaled)
    nF: int
    Rcw: np.ndarray               # (nF,3,3) f64 worldToCam_evalPT rotation
    tcw: np.ndarray               # (nF,3)   f64
    state_zero: np.ndarray        # (nF,10)  f64 unscaled
    state: np.ndarray             # (nF,10)  f64 unscaled
    ab_exposure: np.ndarray       # (nF,)    f32
    frame_id: np.ndarray          # (nF,)    i32 (0 => gauge prior)
    pyramids: list                # [nF][levels] (h_l,w_l,3) f32
    # points (sorted by host), residuals grouped by point
    pt_host: np.ndarray           # (nP,) i32
    pt_u: np.ndarray              # (nP,) f32
    pt_v: np.ndarray
    pt_idepth_zero: np.ndarray
    pt_idepth: np.ndarray
    pt_has_prior: np.ndarray      # (nP,) u8
    pt_color: np.ndarray          # (nP,8) f32
    pt_weights: np.ndarray        # (nP,8) f32
    res_begin: np.ndarray         # (nP+1,) i32 CSR
    res_target: np.ndarray        # (nR,) i32
    pt_idepth_true: np.ndarray = None

    @property
    def nP(self):
        return int(self.pt_host.shape[0])

    @property
    def nR(self):
        return int(self.res_target.shape[0])

    @property
    def res_point(self):
        return np.repeat(np.arange(self.nP, dtype=np.int32), np.diff(self.res_begin)).astype(np.int32)


def make_window(nF=8, pts_per_frame=250, w=640, h=480, seed=42, outlier_frac=0.05,
                K=None, baseline=0.06, levels=None) -> Window:
    """Config-2/3 style window: nF keyframes on a smooth trajectory, every point observed in all
    other keyframes. Frame 0 keeps a = b = 0 (it carries the 1e14 affine gauge prior in the reference)."""
    rng = np.random.default_rng(seed)
    if K is None:
        K = np.array([400.0 * w / 640.0, 400.0 * w / 640.0, (w - 1) / 2.0, (h - 1) / 2.0])
    K = np.asarray(K, dtype=np.float64)
    if levels is None:
        levels = pyr_levels_for(w, h)
    tex = Texture(1234)
    Rcw = np.zeros((nF, 3, 3))
    tcw = np.zeros((nF, 3))
    aff = np.zeros((nF, 2))
    pyramids = []
    for i in range(nF):
        Rwc = so3_exp(0.01 * i * np.array([0.3, 1.0, 0.2]))
        twc = np.array([baseline * i, 0.01 * np.sin(i), 0.02 * i / 0.06 * baseline])
        Rcw[i] = Rwc.T
        tcw[i] = -Rwc.T @ twc
        if i > 0:
            aff[i] = [rng.uniform(-0.05, 0.05), rng.uniform(-5.0, 5.0)]
        img = render(tex, Rcw[i], tcw[i], K, w, h, aff[i, 0], aff[i, 1])
        pyramids.append(make_images(img, levels))

    state_zero = np.zeros((nF, 10))
    state_zero[:, 6] = aff[:, 0] / SCALE_A
    state_zero[:, 7] = aff[:, 1] / SCALE_B
    state = state_zero.copy()
    d_pose = rng.normal(0.0, 2e-3, (nF, 6))
    d_ab = rng.normal(0.0, 1e-3, (nF, 2))
    d_pose[0] = 0.0
    d_ab[0] = 0.0
    state[:, 0:3] += d_pose[:, 0:3] / SCALE_XI_TRANS
    state[:, 3:6] += d_pose[:, 3:6] / SCALE_XI_ROT
    state[:, 6] += d_ab[:, 0] / SCALE_A
    state[:, 7] += d_ab[:, 1] / SCALE_B

    hosts, us, vs, idz, idc, idt, cols, wts = [], [], [], [], [], [], [], []
    for i in range(nF):
        u = rng.integers(20, w - 20, pts_per_frame).astype(np.float64)
        v = rng.integers(20, h - 20, pts_per_frame).astype(np.float64)
        _, _, depth, _ = scene_depth(Rcw[i], tcw[i], K, u, v)
        id_true = 1.0 / depth
        id_zero = id_true * (1.0 + rng.normal(0.0, 0.01, pts_per_frame))
        is_out = rng.uniform(0, 1, pts_per_frame) < outlier_frac
        id_zero = np.where(is_out, id_zero * rng.uniform(0.3, 3.0, pts_per_frame), id_zero)
        id_cur = id_zero + rng.normal(0.0, 5e-3, pts_per_frame)
        id_cur = np.maximum(id_cur, 1e-3)
        c8 = np.zeros((pts_per_frame, 8), np.float32)
        w8 = np.zeros((pts_per_frame, 8), np.float32)
        for k in range(8):
            c, gx, gy = sample_bilin(pyramids[i][0], u + PATTERN[k, 0], v + PATTERN[k, 1])
            c8[:, k] = c
            w8[:, k] = np.sqrt(np.float32(OUTLIER_TH_SUM_COMPONENT) / (np.float32(OUTLIER_TH_SUM_COMPONENT) + (gx * gx + gy * gy)))
        hosts.append(np.full(pts_per_frame, i, np.int32))
        us.append(u.astype(np.float32)); vs.append(v.astype(np.float32))
        idz.append(id_zero.astype(np.float32)); idc.append(id_cur.astype(np.float32)); idt.append(id_true.astype(np.float32))
        cols.append(c8); wts.append(w8)
    pt_host = np.concatenate(hosts)
    nP = pt_host.shape[0]
    targets = []
    for p in range(nP):
        targets.append(np.array([t for t in range(nF) if t != pt_host[p]], np.int32))
    res_begin = np.zeros(nP + 1, np.int32)
    res_begin[1:] = np.cumsum([len(t) for t in targets])
    res_target = np.concatenate(targets) if nP else np.zeros(0, np.int32)
    return Window(w=w, h=h, levels=levels, K=K, nF=nF, Rcw=Rcw, tcw=tcw, state_zero=state_zero, state=state,
                  ab_exposure=np.ones(nF, np.float32), frame_id=np.arange(nF, dtype=np.int32), pyramids=pyramids,
                  pt_host=pt_host, pt_u=np.concatenate(us), pt_v=np.concatenate(vs),
                  pt_idepth_zero=np.concatenate(idz), pt_idepth=np.concatenate(idc),
                  pt_has_prior=np.zeros(nP, np.uint8), pt_color=np.concatenate(cols), pt_weights=np.concatenate(wts),
                  res_begin=res_begin, res_target=res_target, pt_idepth_true=np.concatenate(idt))


def shard_window(win: Window, rank: int, world: int) -> Window:
    """Point-sharded view of a window for rank `rank` of `world`: every rank keeps all frames and
    images and a contiguous slice of each host's points (with their residuals)."""
    keep = np.zeros(win.nP, bool)
    for hst in range(win.nF):
        idx = np.nonzero(win.pt_host == hst)[0]
        n = len(idx)
        lo, hi = (n * rank) // world, (n * (rank + 1)) // world
        keep[idx[lo:hi]] = True
    sel = np.nonzero(keep)[0]
    counts = np.diff(win.res_begin)[sel]
    res_begin = np.zeros(len(sel) + 1, np.int32)
    res_begin[1:] = np.cumsum(counts)
    res_sel = np.concatenate([np.arange(win.res_begin[p], win.res_begin[p + 1]) for p in sel]) if len(sel) else np.zeros(0, np.int64)
    return dataclasses.replace(
        win, pt_host=win.pt_host[sel], pt_u=win.pt_u[sel], pt_v=win.pt_v[sel], pt_idepth_zero=win.pt_idepth_zero[sel],
        pt_idepth=win.pt_idepth[sel], pt_has_prior=win.pt_has_prior[sel], pt_color=win.pt_color[sel],
        pt_weights=win.pt_weights[sel], res_begin=res_begin, res_target=win.res_target[res_sel].astype(np.int32),
        pt_idepth_true=None if win.pt_idepth_true is None else win.pt_idepth_true[sel])


@dataclasses.dataclass
class TrackPair:
    """Config-1 style tracker input: reference keyframe + new frame."""
    w: int
    h: int
    levels: int
    K: np.ndarray
    ref_pyr: list
    new_pyr: list
    ref_aff: np.ndarray           # (2,) a,b of the reference (aff_g2l)
    new_aff_true: np.ndarray
    R_true: np.ndarray            # refToNew
    t_true: np.ndarray
    cpt: np.ndarray               # (n,3) f32 centerProjectedTo of the contributing points (u, v, idepth in ref)
    HdiF: np.ndarray              # (n,)  f32


def make_track_pair(w=640, h=480, n_pts=2000, seed=7, K=None) -> TrackPair:
    rng = np.random.default_rng(seed)
    if K is None:
        K = np.array([400.0 * w / 640.0, 400.0 * w / 640.0, (w - 1) / 2.0, (h - 1) / 2.0])
    K = np.asarray(K, dtype=np.float64)
    levels = pyr_levels_for(w, h)
    tex = Texture(1234)
    R0, t0 = np.eye(3), np.zeros(3)
    axis = np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    R_rn = so3_exp(0.01 * axis)
    t_rn = np.array([0.05, 0.01, 0.02])
    ref_aff = np.array([0.0, 0.0])
    new_aff = np.array([0.02, 1.5])
    ref = render(tex, R0, t0, K, w, h, ref_aff[0], ref_aff[1])
    new = render(tex, R_rn @ R0, R_rn @ t0 + t_rn, K, w, h, new_aff[0], new_aff[1])
    u = rng.integers(20, w - 20, n_pts).astype(np.float64)
    v = rng.integers(20, h - 20, n_pts).astype(np.float64)
    _, _, depth, _ = scene_depth(R0, t0, K, u, v)
    idp = (1.0 / depth) * (1.0 + rng.normal(0.0, 0.02, n_pts))
    cpt = np.stack([u, v, idp], axis=1).astype(np.float32)
    HdiF = rng.uniform(1e-4, 1e-2, n_pts).astype(np.float32)
    return TrackPair(w=w, h=h, levels=levels, K=K, ref_pyr=make_images(ref, levels), new_pyr=make_images(new, levels),
                     ref_aff=ref_aff, new_aff_true=new_aff, R_true=R_rn, t_true=t_rn, cpt=cpt, HdiF=HdiF)


# ---------------------------------------------------------------------------------------------------------
# immature-point tracing: FullSystem::traceNewCoarse's inputs for a synthetic window
@dataclasses.dataclass
class TraceCase:
    w: int
    h: int
    n: int
    u: np.ndarray            # (n,) f32 candidate pixels on their host keyframes (integer positions like the pixel selector's)
    v: np.ndarray
    host: np.ndarray         # (n,) i32 host keyframe index
    KRKi: np.ndarray         # (nF, nF, 3, 3) f32: KRKi[new, host]  (FullSystem.cc:1027-1029)
    Kt: np.ndarray           # (nF, nF, 3) f32
    aff: np.ndarray          # (nF, nF, 2) f32  AffLight::fromToVecExposure(host, new) with exposures 1


def make_trace_case(win: Window, n_per_host=200, seed=5, hosts=None) -> TraceCase:
    rng = np.random.default_rng(seed)
    nF = win.nF
    hosts = list(range(nF - 2)) if hosts is None else list(hosts)
    K32 = np.array([[win.K[0], 0, win.K[2]], [0, win.K[1], win.K[3]], [0, 0, 1]], np.float32)
    Ki32 = np.linalg.inv(K32.astype(np.float64)).astype(np.float32)
    ab = np.stack([win.state_zero[:, 6] * SCALE_A, win.state_zero[:, 7] * SCALE_B], axis=1)
    KRKi = np.zeros((nF, nF, 3, 3), np.float32); Kt = np.zeros((nF, nF, 3), np.float32); aff = np.zeros((nF, nF, 2), np.float32)
    for new in range(nF):
        for h in range(nF):
            R = (win.Rcw[new] @ win.Rcw[h].T)
            t = win.tcw[new] - R @ win.tcw[h]
            KRKi[new, h] = (K32 @ R.astype(np.float32)) @ Ki32
            Kt[new, h] = K32 @ t.astype(np.float32)
            a = np.exp(ab[new, 0] - ab[h, 0])
            aff[new, h] = [a, ab[new, 1] - a * ab[h, 1]]
    us, vs, hs = [], [], []
    for h in hosts:
        us.append(rng.integers(20, win.w - 20, n_per_host).astype(np.float32))
        vs.append(rng.integers(20, win.h - 20, n_per_host).astype(np.float32))
        hs.append(np.full(n_per_host, h, np.int32))
    return TraceCase(w=win.w, h=win.h, n=n_per_host * len(hosts), u=np.concatenate(us), v=np.concatenate(vs), host=np.concatenate(hs),
                     KRKi=KRKi, Kt=Kt, aff=aff)
