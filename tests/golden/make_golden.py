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
    print("golden written")


if __name__ == "__main__":
    main()
