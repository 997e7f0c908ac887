"""Generate the frozen oracle outputs used by tests/test_oracle_cpu.py and the GPU parity tests.

    python tests/golden/make_golden.py
The inputs are the seeded synthetic windows of ldso_b200.synth; the outputs come from oracle/liboracle.so
(-ffp-contract=off build, serial emulation of the reference's 6 worker accumulators).
The reference itself ships no golden vectors for this path and cannot be compiled or imported here.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ldso_b200 import synth  # noqa: E402
from tests import oracle_py  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    win = synth.make_window(nF=5, pts_per_frame=60, w=320, h=240, seed=3)
    o = oracle_py.OracleBA(win, threads_mode=0)
    e0 = o.optimize_begin()
    o.solve_system(0)
    s = o.system()
    r = o.residuals()
    p = o.points()
    P = o.nullspace_projector()
    np.savez_compressed(os.path.join(HERE, "ba_small.npz"), energy0=e0, HA=s["HA"], bA=s["bA"], Hsc=s["Hsc"], bsc=s["bsc"],
                        lastHS=s["lastHS"], lastbS=s["lastbS"], lastX=s["lastX"], P=P, state_NewState=r["state_NewState"],
                        J=r["J"], JpJdF=r["JpJdF"], isActive=r["isActive"], HdiF=p["HdiF"], bdSumF=p["bdSumF"], step=p["step"])
    pair = synth.make_track_pair(w=320, h=240, n_pts=400, seed=7)
    ot = oracle_py.OracleTracker(pair)
    res, H, b = ot.eval(0, np.eye(3), np.zeros(3), 0.0, 0.0, 20.0)
    ok, R, t, a, bb, lr, lf, ne = ot.track(np.eye(3), np.zeros(3), 0.0, 0.0, pair.levels - 1)
    np.savez_compressed(os.path.join(HERE, "tracker_small.npz"), res0=res, H0=H, b0=b, R=R, t=t, aff=np.array([a, bb]), lastResiduals=lr,
                        lastFlow=lf, n_evals=ne)
    # immature-point trace: two traceNewCoarse passes over 150 candidates per host keyframe
    wt = synth.make_window(nF=6, pts_per_frame=10, w=320, h=240, seed=3)
    case = synth.make_trace_case(wt, 150, seed=5)
    tr = oracle_py.OracleTrace(wt, case)
    st1 = tr.trace_on(wt.nF - 2)
    imin1, imax1 = tr.idepth_min.copy(), tr.idepth_max.copy()
    st2 = tr.trace_on(wt.nF - 1)
    np.savez_compressed(os.path.join(HERE, "trace_small.npz"), color=tr.color, weights=tr.weights, gradH=tr.gradH, status1=st1, status2=st2,
                        idepth_min1=imin1, idepth_max1=imax1, idepth_min2=tr.idepth_min, idepth_max2=tr.idepth_max, quality=tr.quality,
                        uv=tr.uv, interval=tr.interval)
    make_select()
    print("golden written")


def select_case():
    """The activation-selection case shared by make_select() and the tests: window, traced candidates, arguments."""
    win = synth.make_window(nF=6, pts_per_frame=40, w=320, h=240, seed=3)
    case = synth.make_trace_case(win, 300, seed=5)
    tr = oracle_py.OracleTrace(win, case)
    tr.trace_on(win.nF - 2)
    tr.trace_on(win.nF - 1)
    newest = win.nF - 1
    m = case.host != newest
    n = int(m.sum())
    my_type = np.random.default_rng(11).choice(np.array([1.0, 2.0, 4.0], np.float32), n)
    quality = np.where(np.isfinite(tr.quality[m]), tr.quality[m], 0).astype(np.float32)
    flagged = np.zeros(win.nF, np.uint8)
    flagged[0] = 1
    args = (case.u[m], case.v[m], case.host[m], tr.idepth_min[m], tr.idepth_max[m], tr.status[m], tr.interval[m], quality, my_type)
    return win, newest, args, flagged


def make_select():
    # activatePointsMT's selection: actions and the level-1 distance map for two minimum distances
    win, newest, args, flagged = select_case()
    o = oracle_py.OracleBA(win, threads_mode=0)
    a13, m13 = o.select_activation(newest, 1.3, *args, frame_flagged=flagged)
    a20, m20 = o.select_activation(newest, 2.0, *args, frame_flagged=flagged)
    _, m0 = o.select_activation(newest, 2.0, *(x[:0] for x in args), frame_flagged=flagged)
    np.savez_compressed(os.path.join(HERE, "select_small.npz"), action13=a13, map13=m13.astype(np.uint16), action20=a20, map20=m20.astype(np.uint16),
                        map_seed_only=m0.astype(np.uint16))


if __name__ == "__main__":
    main()
