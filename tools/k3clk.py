import sys, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldso_b200 import capi, synth
win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
ctx = capi.Context(win.w, win.h, win.levels); ctx.load_synth_window(win)
ctx.optimize_begin(); ctx.gn_iterations(0, 30); ctx.synchronize()
buf = (C.c_longlong*32)()
ctx.L.ldso_b200_debug_clocks(ctx.ctx, buf)
t = np.array(buf[:10]); print("K3 stamps delta cycles:", np.diff(t), "total", t[-1]-t[0])
t = np.array(buf[10:14]); print("K3 block step k0=8 [diag, panel, trailing]:", np.diff(t))
t = np.array(buf[16:24]); print("K1 stamps delta cycles:", np.diff(t), "total", t[-1]-t[0])
t = np.array(buf[24:28]); print("K2b diag-CTA stamps delta cycles:", np.diff(t))
t = np.array(buf[28:30]); print("K2b select-CTA cycles:", np.diff(t))
cap = 1024
sp = (C.c_longlong * (3 * cap))()
n = ctx.L.ldso_b200_debug_cta_spans(ctx.ctx, sp, cap)
a = np.array(sp[:3 * n]).reshape(n, 3)
t0 = a[:, 0].min()
st, en, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2]
print(f"K1 CTA spans (ns, globaltimer): n={n} start p0/p50/p100 = {st.min()}/{int(np.median(st))}/{st.max()}  end p0/p50/p100 = {en.min()}/{int(np.median(en))}/{en.max()}")
dur = en - st
print(f"   duration p0/p50/p90/p100 = {dur.min()}/{int(np.median(dur))}/{int(np.percentile(dur, 90))}/{dur.max()}  distinct SMs {len(set(sm.tolist()))}  max CTAs/SM {np.bincount(sm.astype(int)).max()}")
order = np.argsort(en)[-6:]
print("   last finishers (cta, start, end, dur, sm):", [(int(i), int(st[i]), int(en[i]), int(dur[i]), int(sm[i])) for i in order])
