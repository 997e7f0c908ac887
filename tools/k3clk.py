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
