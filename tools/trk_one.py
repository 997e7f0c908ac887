"""Development aid for ncu: the coarse tracker on BASELINE config 1's pair -- a few level-0 evaluations (k_trk_eval), one whole
trackNewestCoarse (k_trk_track) and one 27-hypothesis batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldso_b200 import capi, synth
from tests import oracle_py
pair = synth.make_track_pair()
ot = oracle_py.OracleTracker(pair, fast=True)
ctx = capi.Context(pair.w, pair.h, pair.levels)
ctx.upload_frame(0, pair.ref_pyr); ctx.upload_frame(1, pair.new_pyr)
ctx.tracker_make_k(*[float(x) for x in pair.K])
for l in range(pair.levels):
    ctx.tracker_set_ref_level(l, *ot.pc(l))
ctx.tracker_set_frames(pair.ref_aff[0], pair.ref_aff[1], 1.0, 1, 1.0)
for _ in range(4):
    ctx.tracker_eval(0, pair.R_true, pair.t_true, 0.0, 0.0, 20.0)
I, z = np.eye(3), np.zeros(3)
for _ in range(3):
    r = ctx.tracker_track(I, z, 0.0, 0.0, pair.levels - 1)
n = 27
rng = np.random.default_rng(1)
Rs = np.stack([synth.so3_exp(rng.normal(0, 0.01, 3)) for _ in range(n)]); Rs[0] = I
b = ctx.tracker_track_batch(Rs, np.zeros((n, 3)), np.zeros((n, 2), np.float32), pair.levels - 1)
print("ok", r[0], int(b["ok"].sum()), "pc_n level 0:", len(ot.pc(0)[0]))
