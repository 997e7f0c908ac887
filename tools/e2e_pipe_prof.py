"""Development aid (GPU box): where does the pipelined end-to-end step (bench.py run_e2e) spend its time? Host wall clock of
ldso_b200_optimize_from_host_submit / _wait per step for (a) two contexts fed alternately from one host thread, (b) three contexts,
(c) two host threads, each driving its own context with the blocking call (ctypes releases the GIL inside the C call)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldso_b200 import capi, synth

win = synth.make_window(nF=8, pts_per_frame=250, seed=42)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
N = 400

def make(n):
    out = []
    for _ in range(n):
        c = capi.Context(win.w, win.h, win.levels, device=0); c.load_synth_window(win)
        out.append(capi.StepIO(c, win, pinned_alloc=pin))
    return out

def pipelined(ios):
    m = len(ios)
    for k in range(2 * m): ios[k % m].fused(0, 1)
    torch.cuda.synchronize()
    ts = tw = 0.0
    t0 = time.perf_counter()
    for k in range(m - 1):
        ios[k].submit(0, 1)
    for k in range(m - 1, N):
        a = time.perf_counter(); ios[k % m].submit(0, 1); b = time.perf_counter(); ios[(k - m + 1) % m].wait(); c = time.perf_counter()
        ts += b - a; tw += c - b
    for k in range(N - m + 1, N): ios[k % m].wait()
    dt = time.perf_counter() - t0
    print(f"{m} contexts, one host thread: {1e6 * dt / N:.1f} us/step ({N / dt:.0f} steps/s); submit {1e6 * ts / (N - m + 1):.1f} us, wait {1e6 * tw / (N - m + 1):.1f} us per step")

def threaded(ios):
    for io in ios: io.fused(0, 1)
    torch.cuda.synchronize()
    per = N // len(ios)
    def work(io):
        for _ in range(per): io.fused(0, 1)
    th = [threading.Thread(target=work, args=(io,)) for io in ios]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{len(ios)} host threads, one context each, blocking call: {1e6 * dt / (per * len(ios)):.1f} us/step ({per * len(ios) / dt:.0f} steps/s)")

ios = make(3)
t0 = time.perf_counter()
for _ in range(100): ios[0].fused(0, 1)
print(f"one context, blocking: {1e4 * (time.perf_counter() - t0):.1f} us/step")
pipelined(ios[:2]); pipelined(ios[:3]); threaded(ios[:2]); threaded(ios[:3])
# host-only cost of the pieces of a submit (async; the stream drains between pieces)
io = ios[0]; L, h = io.L, io.h
io._prep(0, 1)
for name, f in (("make_images", lambda: L.ldso_b200_make_images(h, io.nF - 1, io._color)),
                ("set_frames", lambda: L.ldso_b200_set_frames(h, io.nF, io._frames, io._Ks, io._Kz)),
                ("set_window", lambda: L.ldso_b200_set_window(h, io._wref)),
                ("optimize_begin", lambda: L.ldso_b200_optimize_begin(h, None)),
                ("gn_iterations", lambda: L.ldso_b200_gn_iterations(h, 0, 1)),
                ("prefetch_results", lambda: L.ldso_b200_prefetch_results(h))):
    acc = 0.0
    for _ in range(50):
        torch.cuda.synchronize(); a = time.perf_counter(); f(); acc += time.perf_counter() - a
    print(f"   host time of {name:18s} {1e6 * acc / 50:7.1f} us")
