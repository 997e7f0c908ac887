"""Development aid: how far are the device trace results from the oracle's, bit for bit? Prints, per geometry and pass, the number of
candidates whose status / interval / position / quality differ at all, and the first few differing candidates with both sides' values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldso_b200 import capi, synth
from tests import oracle_py
from tests.test_gpu_trace import _fresh

for geom in ("small", "vga", "kitti"):
    if geom == "small":
        win = synth.make_window(nF=6, pts_per_frame=10, w=320, h=240, seed=3); per_host = 150
    elif geom == "vga":
        win = synth.make_window(nF=8, pts_per_frame=10, seed=42); per_host = 250
    else:
        win = synth.make_window(nF=5, pts_per_frame=10, w=1232, h=368, seed=11, K=np.array([718.856, 718.856, 607.1928, 185.2157])); per_host = 300
    case = synth.make_trace_case(win, per_host, seed=5)
    ctx = capi.Context(win.w, win.h, win.levels)
    for i in range(win.nF):
        ctx.upload_frame(i, win.pyramids[i])
    tr = oracle_py.OracleTrace(win, case)
    pts = _fresh(case, dict(color=tr.color, weights=tr.weights, gradH=tr.gradH, energyTH=tr.energyTH))
    for new in (win.nF - 2, win.nF - 1):
        so = tr.trace_on(new).copy()
        ctx.trace_immature(new, pts, case.KRKi[new], case.Kt[new], case.aff[new])
        d = {"status": pts["status"] != so}
        for k, o in (("idepth_min", tr.idepth_min), ("idepth_max", tr.idepth_max), ("quality", tr.quality), ("interval", tr.interval)):
            d[k] = ~((pts[k] == o) | (np.isnan(pts[k]) & np.isnan(o)))
        d["uv"] = np.any(~((pts["uv"] == tr.uv) | (np.isnan(pts["uv"]) & np.isnan(tr.uv))), axis=1)
        print(geom, "new", new, "n", case.n, {k: int(v.sum()) for k, v in d.items()})
        anyd = np.nonzero(np.any(np.stack(list(d.values())), axis=0))[0]
        for i in anyd[:6]:
            print("   cand", i, "host", case.host[i], "status", pts["status"][i], so[i], "idmin", pts["idepth_min"][i], tr.idepth_min[i],
                  "idmax", pts["idepth_max"][i], tr.idepth_max[i], "q", pts["quality"][i], tr.quality[i], "uv", pts["uv"][i], tr.uv[i],
                  "int", pts["interval"][i], tr.interval[i])
        for k, o in (("idepth_min", tr.idepth_min), ("idepth_max", tr.idepth_max), ("quality", tr.quality), ("status", tr.status), ("interval", tr.interval)):
            pts[k][:] = o
        pts["uv"][:] = tr.uv
    ctx.close()
