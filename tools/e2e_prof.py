"""Per-call host wall-clock breakdown of bench.py's end-to-end step (development aid, GPU box only)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldso_b200 import capi, synth

win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = capi.Context(win.w, win.h, win.levels, device=0); ctx.set_stream(stream.cuda_stream)
ctx.load_synth_window(win)
nF = win.nF
color = np.ascontiguousarray(win.pyramids[nF - 1][0][:, :, 0])
pin = torch.from_numpy(color).pin_memory().numpy()
calls = [
    ("make_images", lambda: ctx.make_images(nF - 1, pin)),
    ("set_frames", lambda: ctx.set_frames(win.Rcw, win.tcw, win.state_zero, win.state, win.ab_exposure, win.frame_id, list(range(nF)), win.K)),
    ("set_window", lambda: ctx.set_window(win.pt_host, win.pt_u, win.pt_v, win.pt_idepth, win.pt_idepth_zero, win.pt_has_prior, win.pt_color, win.pt_weights, win.res_begin, win.res_target)),
    ("optimize_begin", lambda: ctx.optimize_begin(want_energy=False)),
    ("gn_iterations", lambda: ctx.gn_iterations(0, 1)),
    ("last_solution", lambda: ctx.last_solution()),
    ("points", lambda: ctx.points()),
    ("residuals", lambda: ctx.residuals(with_J=False)),
]
for mode in ("sync_after_each", "async"):
    acc = {k: 0.0 for k, _ in calls}; accs = {k: 0.0 for k, _ in calls}
    N = 60
    for it in range(N + 5):
        for k, f in calls:
            t0 = time.perf_counter(); f(); t1 = time.perf_counter()
            if mode == "sync_after_each": torch.cuda.synchronize()
            t2 = time.perf_counter()
            if it >= 5: acc[k] += t1 - t0; accs[k] += t2 - t1
    print(mode, "total us/step", round(1e6 * sum(acc.values()) / N + 1e6 * sum(accs.values()) / N, 1))
    for k, _ in calls: print(f"  {k:16s} call {1e6*acc[k]/N:8.1f} us   sync-tail {1e6*accs[k]/N:8.1f} us")

io = capi.StepIO(ctx, win, pinned_alloc=lambda a: torch.from_numpy(a).pin_memory().numpy())
for mode in ("sync_after_each", "async"):
    names = ("upload", "step", "download"); fs = (io.upload, io.step, io.download)
    acc = dict.fromkeys(names, 0.0); accs = dict.fromkeys(names, 0.0)
    N = 100
    for it in range(N + 5):
        for k, f in zip(names, fs):
            t0 = time.perf_counter(); f(); t1 = time.perf_counter()
            if mode == "sync_after_each": torch.cuda.synchronize()
            t2 = time.perf_counter()
            if it >= 5: acc[k] += t1 - t0; accs[k] += t2 - t1
    print("StepIO", mode, "total us/step", round(1e6 * (sum(acc.values()) + sum(accs.values())) / N, 1))
    for k in names: print(f"  {k:16s} call {1e6*acc[k]/N:8.1f} us   sync-tail {1e6*accs[k]/N:8.1f} us")
# finer: individual raw calls
L, h = io.L, io.h
raw = [("make_images", lambda: L.ldso_b200_make_images(h, io.nF - 1, io._color)),
       ("set_frames", lambda: L.ldso_b200_set_frames(h, io.nF, io._frames, io._Ks, io._Kz)),
       ("set_window", lambda: L.ldso_b200_set_window(h, io._wref)),
       ("optimize_begin", lambda: L.ldso_b200_optimize_begin(h, None)),
       ("gn_iterations", lambda: L.ldso_b200_gn_iterations(h, 0, 1)),
       ("prefetch", lambda: L.ldso_b200_prefetch_results(h)),
       ("get_last_solution", lambda: L.ldso_b200_get_last_solution(h, *io._sol)),
       ("get_points", lambda: L.ldso_b200_get_points(h, *io._pts)),
       ("get_residuals", lambda: L.ldso_b200_get_residuals(h, *io._res))]
acc = {k: 0.0 for k, _ in raw}
N = 100
for it in range(N + 5):
    for k, f in raw:
        t0 = time.perf_counter(); f(); t1 = time.perf_counter()
        if it >= 5: acc[k] += t1 - t0
print("raw async total us/step", round(1e6 * sum(acc.values()) / N, 1))
for k, _ in raw: print(f"  {k:18s} {1e6*acc[k]/N:8.1f} us")
