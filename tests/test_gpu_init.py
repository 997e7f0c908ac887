"""CoarseInitializer::calcResAndGS on the device against the oracle, which oracle/ref_pin pins bit for bit against the
reference's own CoarseInitializer.cc. Per-point outputs (decisions, energies, maxstep, JbBuffer) are compared bit for bit (the kernel's
translation unit is built with -fmad=false and folds the pattern in the reference's order); the 45 + 45 Hessian sums and the energy to
1e-4 (their cross-point summation order differs from the reference's SSE lanes). The oracle half runs on the CPU."""
import os

import numpy as np
import pytest

from ldso_b200 import synth
from tests import oracle_py


def _case(pair, lvl, seed):
    rng = np.random.default_rng(seed)
    wl, hl = pair.w >> lvl, pair.h >> lvl
    n = 3000 >> lvl
    u = rng.integers(3, wl - 4, n).astype(np.float32); v = rng.integers(3, hl - 4, n).astype(np.float32)
    idn = rng.uniform(0.4, 2.2, n).astype(np.float32); iR = rng.uniform(0.8, 1.2, n).astype(np.float32)
    good = (rng.uniform(size=n) > 0.1).astype(np.uint8)
    e2 = np.stack([rng.uniform(0, 300, n), rng.uniform(0, 1, n)], 1).astype(np.float32)
    oth = np.where(rng.uniform(size=n) < 0.08, 0.5, 8 * 12 * 12.0).astype(np.float32)
    return u, v, idn, iR, good, e2, oth


def test_init_calc_res_oracle():
    """The oracle alone (CPU): identical frames and the identity pose give a zero photometric energy; a real pair gives a symmetric,
    positive semi-definite H and rejects the points whose pattern leaves the image."""
    pair = synth.make_track_pair(w=320, h=240, n_pts=400, seed=7)
    u, v, idn, iR, good, e2, oth = _case(pair, 0, 1)
    o = oracle_py.init_calc_res(pair, 0, pair.R_true, pair.t_true, 0.0, 0.0, u, v, idn, iR, good, e2, oth)
    assert np.allclose(o["H"], o["H"].T) and np.linalg.eigvalsh(o["H"].astype(np.float64)).min() > -1e-3 * np.abs(o["H"]).max()
    assert (o["isGood_new"][good == 0] == 0).all() and o["isGood_new"].sum() > 0.3 * len(u)
    assert (o["maxstep"][good == 0] == np.float32(1e10)).all()
    assert o["res"][2] == 2 * len(u)


@pytest.mark.gpu
@pytest.mark.parametrize("lvl", [0, 2])
def test_init_calc_res_matches_oracle(lvl):
    from ldso_b200 import capi
    pair = synth.make_track_pair()
    u, v, idn, iR, good, e2, oth = _case(pair, lvl, 3 + lvl)
    ctx = capi.Context(pair.w, pair.h, pair.levels)
    ctx.upload_frame(0, pair.ref_pyr)
    ctx.upload_frame(1, pair.new_pyr)
    for R, t, a, b in ((np.eye(3), np.zeros(3), 0.0, 0.0), (pair.R_true, pair.t_true, 0.01, 1.0), (pair.R_true, 40 * pair.t_true, 0.0, 0.0)):
        tl = oracle_py.se3_log(R, t)[:3]
        o = oracle_py.init_calc_res(pair, lvl, R, t, a, b, u, v, idn, iR, good, e2, oth)
        g = ctx.init_calc_res(0, 1, lvl, R, t, tl, a, b, pair.K, u, v, idn, iR, good, e2, oth)
        acc = o["isGood_new"] == 1
        live = good == 1
        bad = {"isGood_new": int(np.sum(g["isGood_new"] != o["isGood_new"])),
               "energy_new": int(np.sum(g["energy_new"] != o["energy_new"])),
               "maxstep": int(np.sum(g["maxstep"] != o["maxstep"])),
               "lastHessian_new": int(np.sum(g["lastHessian_new"][acc] != o["lastHessian_new"][acc])),
               "Jb": int(np.sum(g["Jb"][live] != o["Jb"][live]))}
        assert not any(bad.values()), f"bit mismatches (entries): {bad}"
        for k in ("H", "b", "Hsc", "bsc"):
            assert np.linalg.norm(g[k].astype(np.float64) - o[k]) <= 1e-4 * np.linalg.norm(o[k]) + 1e-6, k
        assert abs(g["res"][0] - o["res"][0]) <= 1e-4 * abs(o["res"][0]) and g["res"][1] == o["res"][1] and g["res"][2] == o["res"][2]
    ctx.close()
