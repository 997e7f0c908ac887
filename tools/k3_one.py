"""Development aid for ncu: a few GN iterations on the bench window with plain launches (LDSO_B200_NO_GRAPH=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldso_b200 import capi, synth
pts = int(sys.argv[1]) if len(sys.argv) > 1 else 250
win = synth.make_window(nF=8, pts_per_frame=pts, seed=42)
ctx = capi.Context(win.w, win.h, win.levels); ctx.load_synth_window(win)
ctx.optimize_begin(); ctx.gn_iterations(0, 6); ctx.synchronize()
print("ok", ctx.energy())
