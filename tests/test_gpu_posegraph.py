"""Sim(3) pose-graph optimisation (BASELINE configs[4]) on the device against the numpy / scipy oracle
(oracle/posegraph.py: Map.cc:75-165 + g2o's numeric-Jacobian Gauss-Newton + Sophus' Sim3 restated; parity unpinned, see its header).
Numeric differentiation with delta = 1e-9 carries ~1e-7 relative noise in every Jacobian entry ON BOTH SIDES (the reference's own
g2o run included), so the trajectories are compared to 1e-5 (chi2) / 1e-6 (poses), not bit for bit."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldso_b200 import capi
from oracle import posegraph as pg


def test_oracle_posegraph_cpu():
    """The oracle alone: exp / log round trip, zero error at the ground truth, Gauss-Newton recovers the ground truth."""
    rng = np.random.default_rng(0)
    a = rng.normal(0, 0.5, (500, 7)); a[:, 6] *= 0.2
    assert np.abs(pg.sim3_log(pg.sim3_exp(a)) - a).max() < 1e-10
    g = pg.make_graph(120, 200, seed=2)
    assert pg.linearize(g["gq"], g["gt"], g["ei"], g["ej"], g["mq"], g["mt"], g["info"])[3] < 1e-20
    q, t, chi = pg.optimize(g["q"], g["t"], g["ei"], g["ej"], g["mq"], g["mt"], g["info"], g["fixed"], iterations=6)
    assert chi[0] > 1e-2 and chi[-1] < 1e-12 * chi[0]
    assert np.abs(t - g["gt"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(150, 250, 0.0), (600, 1200, 0.02)])
def test_posegraph_matches_oracle(size):
    n_kf, n_loop, meas_noise = size
    g = pg.make_graph(n_kf, n_loop, seed=5)
    if meas_noise > 0:      # inconsistent measurements: a non-zero optimum, several Gauss-Newton rounds with a changing system
        rng = np.random.default_rng(9)
        pert = rng.normal(0, meas_noise, (len(g["ei"]), 7)); pert[:, 3:6] *= 0.2; pert[:, 6] *= 0.1
        g["mq"], g["mt"] = pg.sim3_mul(pg.sim3_exp(pert), (g["mq"], g["mt"]))
        rng2 = np.random.default_rng(10)
        A = rng2.normal(0, 0.3, (len(g["ei"]), 7, 7))
        g["info"] = np.eye(7) + A @ np.transpose(A, (0, 2, 1))          # full (SPD) information matrices
    its = 6
    qo, to, co = pg.optimize(g["q"], g["t"], g["ei"], g["ej"], g["mq"], g["mt"], g["info"], g["fixed"], iterations=its)
    ctx = capi.Context(64, 64, 1)
    qg, tg, cg, ncg = ctx.posegraph_optimize(g["q"], g["t"], g["ei"], g["ej"], g["mq"], g["mt"], g["info"], g["fixed"], iterations=its, pcg_tol=1e-12)
    ctx.close()
    assert abs(cg[0] - co[0]) <= 1e-9 * co[0]                                   # same linearisation point: chi2 to rounding
    # later rounds: the numeric Jacobians' 1e-7 noise moves each step by ~1e-7 of its length, i.e. the chi2 that is LEFT after a step
    # that removed five orders of magnitude by a relative 1e-5 or so: the bar is relative to the current value plus 1e-8 of the start
    assert np.all(np.abs(cg - co) <= 1e-5 * co + 1e-8 * co[0]), (cg, co)
    assert np.array_equal(qg[g["fixed"]], g["q"][g["fixed"]]) and np.array_equal(tg[g["fixed"]], g["t"][g["fixed"]])
    sgn = np.sign(np.sum(qg * qo, -1, keepdims=True))                            # q and -q are the same rotation
    assert np.abs(qg * sgn - qo).max() < 1e-6 and np.abs(tg - to).max() < 1e-6 * max(1.0, np.abs(to).max())
    assert ncg > 0
