"""The C++ host shim (ldso_b200/host/ldso_shim.hpp: EnergyFunctional / PointFrameResidual / CoarseTracker with the reference's
signatures) driven like FullSystem::optimize drives the reference classes, compared with the oracle."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from ldso_b200 import synth
from tests import oracle_py, shim_io
from tests.parity import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def win():
    return synth.make_window(nF=5, pts_per_frame=60, w=320, h=240, seed=3)


def _run(win, *args):
    exe = shim_io.build()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "win.bin")
        shim_io.dump(win, path)
        r = subprocess.run([exe, path, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout)


def test_piecewise_loop_like_fullsystem(win):
    out = _run(win)
    o = oracle_py.OracleBA(win, threads_mode=0)
    e = [o.optimize_begin()]
    o.solve_system(0)
    x0 = o.system()["lastX"].copy()
    P = o.nullspace_projector()
    o.do_step(); o.linearize_all(False); o.apply_res()
    e.append(o.L.oracle_ba_last_energy(o.o))
    for it in (1, 2):
        o.gn_iteration(it)
        e.append(o.L.oracle_ba_last_energy(o.o))
    assert np.allclose(out["energy"], e, rtol=2e-3)
    assert abs(out["energy"][0] - e[0]) <= 1e-5 * e[0]
    I = np.eye(P.shape[0])
    assert rel_err((I - P) @ np.array(out["lastX0"]), (I - P) @ x0) < 1e-4
    ro = o.residuals()
    assert abs(out["nIn"] - int((ro["state_state"] == 0).sum())) <= 2
    assert abs(out["nActive"] - int(ro["isActive"].sum())) <= 2
    assert rel_err(out["idepth"], o.points()["idepth"]) < 2e-3
    assert out["track_ok"] is True
    # newest keyframe -> previous keyframe: the tracker must recover the (known) relative translation direction
    Tn = np.eye(4); Tn[:3, :3] = win.Rcw[-1]; Tn[:3, 3] = win.tcw[-1]
    Tp = np.eye(4); Tp[:3, :3] = win.Rcw[-2]; Tp[:3, 3] = win.tcw[-2]
    rel = Tp @ np.linalg.inv(Tn)
    t = np.array(out["track_t"])
    assert np.linalg.norm(t - rel[:3, 3]) < 0.35 * np.linalg.norm(rel[:3, 3])


def test_fused_entry_point(win):
    out = _run(win, "fused")
    assert out["ok"] is True
    o = oracle_py.OracleBA(win, threads_mode=0)
    o.optimize_begin()
    for it in range(3):
        o.gn_iteration(it)
    assert abs(out["energy"][0] - o.L.oracle_ba_last_energy(o.o)) <= 2e-3 * out["energy"][0]
    assert rel_err(out["idepth"], o.points()["idepth"]) < 2e-3


def test_marginalize_frame_through_shim(win):
    """EnergyFunctional::marginalizeFrame with the reference's signature (shim) on a synthetic prior against the oracle."""
    out = _run(win, "margframe")
    assert out["marg_ok"] is True and out["marg_nframes"] == win.nF - 1
    n = 8 * win.nF + 4
    i = np.arange(n, dtype=np.float64)
    v = 20.0 * np.sin(i + 1.0)
    HM = np.outer(v, v) + np.diag(50.0 + 3.0 * i)
    bM = 10.0 * np.cos(i)
    o = oracle_py.OracleBA(win, threads_mode=0)
    o.optimize_begin()
    for it in range(3):
        o.gn_iteration(it)              # same states as the shim after its fused 3 iterations
    o.set_marg_prior(HM, bM)
    Ho, bo = o.marginalize_frame(1)
    nd = out["marg_n"]
    assert nd == n - 8
    Hg = np.array(out["HM"]).reshape(nd, nd, order="F")
    assert rel_err(Hg, Ho) < 1e-9            # HM does not depend on the states
    assert rel_err(np.array(out["bM"]), bo) < 2e-3   # bM carries prior * delta_prior of the 3-iteration states (same bar as idepth above)
