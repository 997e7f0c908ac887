"""Development aid: print the actual device-vs-oracle differences behind the tolerances of tests/ (tracker LM loop at the three
geometries), so that the asserts can be set from measured numbers instead of guesses."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ldso_b200 import capi, synth
from tests import oracle_py
from tests.parity import rel_err
from tests.test_gpu_tracker import _setup

for name, pair in (("small", synth.make_track_pair(w=320, h=240, n_pts=400, seed=7)), ("cfg1", synth.make_track_pair()),
                   ("kitti", synth.make_track_pair(w=1232, h=368, n_pts=1500, seed=9, K=np.array([718.856, 718.856, 607.1928, 185.2157])))):
    ot, ctx = _setup(pair)
    okg, Rg, tg, ag, bg, lrg, lfg = ctx.tracker_track(np.eye(3), np.zeros(3), 0.0, 0.0, pair.levels - 1)
    oko, Ro, to, ao, bo, lro, lfo, ne = ot.track(np.eye(3), np.zeros(3), 0.0, 0.0, pair.levels - 1)
    print(name, "ok", okg, oko, "R", rel_err(Rg, Ro), "t", rel_err(tg, to), "a", abs(ag - ao), "b", abs(bg - bo),
          "lastRes", rel_err(np.nan_to_num(lrg), np.nan_to_num(lro)), "flow", rel_err(lfg, lfo), "evals", ne,
          "t_vs_truth", np.linalg.norm(tg - pair.t_true) / np.linalg.norm(pair.t_true), "oracle t_vs_truth", np.linalg.norm(to - pair.t_true) / np.linalg.norm(pair.t_true))
    ctx.close()
