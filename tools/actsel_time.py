"""Development aid: time ldso_b200_select_activation on the bench window (run with LDSO_B200_KTIME=1 for kernel time and phase cycles)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from ldso_b200 import capi, synth

win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
ctx = capi.Context(win.w, win.h, win.levels)
ctx.load_synth_window(win)
for rep in range(2):
    r = bench.run_select(ctx, win)
    print({k: r[k] for k in ("candidates", "ms_per_call", "cpu_port_ms_per_call_1core", "selected", "identical_to_cpu_port")})
ctx.close()
