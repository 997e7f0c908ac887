"""Development aid (GPU box): cold-L2 time per Gauss-Newton iteration of the headline window for the library LDSO_B200_LIB points at
(default: the in-tree build). Same timing rules as bench.py (192 MB L2 flush before every step, CUDA events on the launching stream).
    LDSO_B200_LIB=ldso_b200/lib/libldso_b200_base.so python tools/ab_time.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldso_b200 import capi, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = capi.Context(win.w, win.h, win.levels, device=0); ctx.set_stream(stream.cuda_stream)
ctx.load_synth_window(win)
ctx.optimize_begin(want_energy=False)
for i in range(10):
    ctx.gn_iterations(min(i, 3), 1)
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for k in range(steps):
    flush.fill_(k & 0xff)
    ev[k][0].record(stream); ctx.gn_iterations(3, 1); ev[k][1].record(stream)
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
ev2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
ev2[0].record(stream); ctx.gn_iterations(3, 200); ev2[1].record(stream); torch.cuda.synchronize()
warm = ev2[0].elapsed_time(ev2[1]) * 1e3 / 200
ctx.kernel_times(True)
for k in range(40):
    flush.fill_(k & 0xff); ctx.gn_iterations(3, 1)
kt = ctx.kernel_times(False)
sol = ctx.last_solution()
print(f"{os.environ.get('LDSO_B200_LIB', 'in-tree')}: cold us/iter mean {t.mean():.2f} median {np.median(t):.2f} p10 {np.percentile(t, 10):.2f} | warm {warm:.2f} | "
      f"plain-launch kernel us { {k: round(v, 1) for k, v in kt.items()} } | |lastX| {np.linalg.norm(sol['lastX']):.9e} energy {ctx.energy()[0]:.6f}")
ctx.close()
