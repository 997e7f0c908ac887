"""Stage-by-stage GPU-vs-oracle comparison (development aid; run on the GPU box).

    python tools/gpu_diag.py [--small] [--out gpurun_out/diag.txt]
Prints one line per compared quantity with the norm-relative error; never asserts, so one run shows everything.
"""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ldso_b200 import capi, synth  # noqa: E402
from tests import oracle_py  # noqa: E402
from tests.parity import rel_err, max_rel  # noqa: E402

LINES = []


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LINES.append(s)


def cmp(name, g, o, kind="rel"):
    try:
        e = rel_err(g, o) if kind == "rel" else max_rel(g, o)
        flag = "" if e < 1e-4 else "   <<<<<< FAIL"
        say(f"  {name:34s} {kind}err={e:.3e}  |ref|={np.linalg.norm(np.asarray(o, np.float64)):.4e}{flag}")
    except Exception as ex:  # noqa: BLE001
        say(f"  {name}: compare failed: {ex}")


def compare_residuals(rg, ro, tag):
    n = len(ro["state_NewState"])
    mism = int(np.sum(rg["state_NewState"].astype(int) != ro["state_NewState"].astype(int)))
    say(f"  [{tag}] NewState mismatches: {mism} / {n}   gpu={np.bincount(rg['state_NewState'], minlength=3)} ref={np.bincount(ro['state_NewState'], minlength=3)}")
    same = rg["state_NewState"].astype(int) == ro["state_NewState"].astype(int)
    notoob = same & (ro["state_NewState"] != 1)
    cmp(f"[{tag}] NewEnergy", rg["state_NewEnergy"][same], ro["state_NewEnergy"][same])
    cmp(f"[{tag}] NewEnergyWithOutlier", rg["state_NewEnergyWithOutlier"][notoob], ro["state_NewEnergyWithOutlier"][notoob])
    if "J" in rg:
        names = [("resF", 0, 8), ("Jpdxi", 8, 20), ("Jpdc", 20, 28), ("Jpdd", 28, 30), ("JIdx", 30, 46), ("JabF", 46, 62),
                 ("JIdx2", 62, 66), ("JabJIdx", 66, 70), ("Jab2", 70, 74)]
        for nm, a, b in names:
            cmp(f"[{tag}] J.{nm}", rg["J"][notoob][:, a:b], ro["J"][notoob][:, a:b])
        cmp(f"[{tag}] projectedTo", rg["projectedTo"][notoob], ro["projectedTo"][notoob])
        cmp(f"[{tag}] centerProjectedTo", rg["centerProjectedTo"][notoob], ro["centerProjectedTo"][notoob])


def run_ba(win, tag):
    say(f"==== BA window {tag}: nF={win.nF} nP={win.nP} nR={win.nR} {win.w}x{win.h}")
    o = oracle_py.OracleBA(win, threads_mode=0)
    ctx = capi.Context(win.w, win.h, win.levels)
    ctx.load_synth_window(win)
    fo, fg = o.frames(), ctx.frames()
    for k in ("precalc", "adHost", "adTarget", "adHTdeltaF", "state", "calib_value"):
        cmp(f"frames.{k}", fg[k], fo[k])
    cmp("nullspace projector", ctx.nullspace_projector(), o.nullspace_projector())

    # ---- piecewise path
    say("-- piecewise: linearize_all / apply_res / solve_system / do_step")
    o.optimize_begin()          # resetOOB + linearizeAll(false) + applyRes
    eg = ctx.linearize_all(False)
    say(f"  energy gpu={eg:.6f} ref={o.L.oracle_ba_last_energy(o.o):.6f}")
    rg, ro = ctx.residuals(), o.residuals()
    compare_residuals(rg, ro, "lin0")
    cmp("frameEnergyTH", ctx.frames()["frameEnergyTH"], o.frames()["frameEnergyTH"])
    ctx.apply_res()
    rg = ctx.residuals()
    say(f"  isActive mismatches: {int(np.sum(rg['isActive'] != ro['isActive']))}")
    cmp("JpJdF (active)", rg["JpJdF"][ro["isActive"] == 1], ro["JpJdF"][ro["isActive"] == 1])
    for it in range(3):
        ctx.backup_state()
        HS, bS, X = ctx.solve_system(it)
        o.solve_system(it)
        so, sg = o.system(), ctx.system()
        say(f"  -- solve iteration {it}: resInA gpu={sg['resInA']} ref={o.res_counts()[0]}")
        for k in ("HA", "bA", "Hsc", "bsc"):
            cmp(f"it{it} {k}", sg[k], so[k])
        cmp(f"it{it} lastHS", HS, so["lastHS"])
        cmp(f"it{it} lastbS", bS, so["lastbS"])
        cmp(f"it{it} lastX", X, so["lastX"])
        Pn = o.nullspace_projector()
        I = np.eye(Pn.shape[0])
        cmp(f"it{it} lastX gauge-projected", (I - Pn) @ X, (I - Pn) @ so["lastX"])
        # the device solver against a double-precision numpy solve of ITS OWN system (isolates K3 from input noise)
        lam = 1e-5
        Hf = HS + sg["Hsc"]
        H2 = Hf.copy(); H2[np.diag_indices_from(H2)] *= (1 + lam); H2 -= sg["Hsc"] / (1 + lam)
        Sv = 1 / np.sqrt(np.diag(H2) + 10)
        xn = Sv * np.linalg.solve(Sv[:, None] * H2 * Sv[None, :], Sv * bS)
        if it >= 2:
            xn = xn - ctx.nullspace_projector() @ xn
        cmp(f"it{it} lastX vs numpy(own system)", X, xn)
        cmp(f"it{it} lastX vs numpy, gauge-projected", (I - Pn) @ X, (I - Pn) @ xn)
        pg, po = ctx.points(), o.points()
        for k in ("HdiF", "bdSumF", "Hcd_accAF", "Hdd_accAF", "bd_accAF"):
            cmp(f"it{it} pt.{k}", pg[k], po[k])
        cmp(f"it{it} pt.step", pg["step"], po["step"], "max")
        cbg = ctx.do_step()
        cbo = o.do_step()
        say(f"  canbreak gpu={cbg} ref={cbo}")
        fo, fg = o.frames(), ctx.frames()
        for k in ("state", "precalc", "adHTdeltaF", "calib_value"):
            cmp(f"it{it} frames.{k}", fg[k], fo[k])
        cmp(f"it{it} pt.idepth", ctx.points()["idepth"], o.points()["idepth"])
        eo = o.linearize_all(False)
        eg = ctx.linearize_all(False)
        say(f"  energy gpu={eg:.6f} ref={eo:.6f} rel={abs(eg - eo) / abs(eo):.2e}")
        compare_residuals(ctx.residuals(), o.residuals(), f"lin{it + 1}")
        cmp(f"it{it} frameEnergyTH", ctx.frames()["frameEnergyTH"], o.frames()["frameEnergyTH"])
        o.apply_res()
        ctx.apply_res()
    ctx.close()

    # ---- fused path
    say("-- fused: optimize_begin + gn_iterations")
    o = oracle_py.OracleBA(win, threads_mode=0)
    ctx = capi.Context(win.w, win.h, win.levels)
    ctx.load_synth_window(win)
    eo = o.optimize_begin()
    eg = ctx.optimize_begin()
    say(f"  begin energy gpu={eg:.6f} ref={eo:.6f} rel={abs(eg - eo) / abs(eo):.2e}")
    for it in range(5):
        ctx.gn_iterations(it, 1)
        ctx.synchronize()
        o.gn_iteration(it)
        sol, so = ctx.last_solution(), o.system()
        for k in ("lastHS", "lastbS", "lastX"):
            cmp(f"fused it{it} {k}", sol[k], so[k])
        e, cb = ctx.energy()
        eo = o.L.oracle_ba_last_energy(o.o)
        say(f"  fused it{it} energy gpu={e:.6f} ref={eo:.6f} rel={abs(e - eo) / abs(eo):.2e} canbreak={cb}")
        cmp(f"fused it{it} pt.idepth", ctx.points()["idepth"], o.points()["idepth"])
        cmp(f"fused it{it} frames.state", ctx.frames()["state"], o.frames()["state"])
        rg, ro = ctx.residuals(with_J=False), o.residuals()
        say(f"  fused it{it} state mismatches: {int(np.sum(rg['state_state'].astype(int) != ro['state_state'].astype(int)))} "
            f"active mismatches: {int(np.sum(rg['isActive'] != ro['isActive']))}")
    say(f"  launches: {ctx.launch_count()}")
    # timing of the fused loop
    ctx.gn_iterations(5, 20)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.gn_iterations(25, 200)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 200
    say(f"  fused GN iteration wall: {dt * 1e6:.1f} us  ({1 / dt:.0f} it/s)")
    ctx.close()


def run_images(win):
    say("==== make_images")
    ctx = capi.Context(win.w, win.h, win.levels)
    color = win.pyramids[1][0][:, :, 0]
    ctx.make_images(3, color)
    for l in range(win.levels):
        cmp(f"makeImages lvl{l}", ctx.download_frame_level(3, l), win.pyramids[1][l])
    ctx.upload_frame(2, win.pyramids[0])
    cmp("upload/download lvl0", ctx.download_frame_level(2, 0), win.pyramids[0][0])
    ctx.close()


def run_tracker(pair):
    say(f"==== tracker {pair.w}x{pair.h} levels={pair.levels}")
    ot = oracle_py.OracleTracker(pair)
    ctx = capi.Context(pair.w, pair.h, pair.levels)
    ctx.upload_frame(0, pair.ref_pyr)
    ctx.upload_frame(1, pair.new_pyr)
    ctx.tracker_make_k(*[float(x) for x in pair.K])
    for l in range(pair.levels):
        u, v, idp, col = ot.pc(l)
        say(f"  lvl {l}: pc_n={len(u)}")
        ctx.tracker_set_ref_level(l, u, v, idp, col)
    ctx.tracker_set_frames(pair.ref_aff[0], pair.ref_aff[1], 1.0, 1, 1.0)
    R0, t0 = np.eye(3), np.zeros(3)
    for l in range(pair.levels - 1, -1, -1):
        for (R, t, tag) in ((R0, t0, "identity"), (pair.R_true, pair.t_true, "truth")):
            rg, Hg, bg = ctx.tracker_eval(l, R, t, 0.0, 0.0, 20.0)
            ro, Ho, bo = ot.eval(l, R, t, 0.0, 0.0, 20.0)
            cmp(f"lvl{l} {tag} res6", rg, ro)
            cmp(f"lvl{l} {tag} H", Hg, Ho)
            cmp(f"lvl{l} {tag} b", bg, bo)
    okg, Rg, tg, ag, bg_, lrg, lfg = ctx.tracker_track(R0, t0, 0.0, 0.0, pair.levels - 1)
    oko, Ro, to, ao, bo_, lro, lfo, ne = ot.track(R0, t0, 0.0, 0.0, pair.levels - 1)
    say(f"  track ok gpu={okg} ref={oko} evals(ref)={ne}")
    cmp("track R", Rg, Ro)
    cmp("track t", tg, to)
    say(f"  aff gpu=({ag:.5f},{bg_:.4f}) ref=({ao:.5f},{bo_:.4f}) truth t={pair.t_true} got t={tg}")
    cmp("track lastResiduals", np.nan_to_num(lrg), np.nan_to_num(lro))
    cmp("track lastFlowIndicators", lfg, lfo)
    t0_ = time.perf_counter()
    for _ in range(20):
        ctx.tracker_track(R0, t0, 0.0, 0.0, pair.levels - 1)
    say(f"  track wall (incl. sync+copies): {(time.perf_counter() - t0_) / 20 * 1e6:.1f} us")
    ctx.close()


def main():
    small = "--small" in sys.argv
    out = "gpurun_out/diag.txt"
    if "--out" in sys.argv:
        out = sys.argv[sys.argv.index("--out") + 1]
    try:
        w1 = synth.make_window(nF=5, pts_per_frame=60, w=320, h=240, seed=3)
        run_images(w1)
        run_ba(w1, "small")
        run_tracker(synth.make_track_pair(w=320, h=240, n_pts=400, seed=7))
        if not small:
            run_ba(synth.make_window(nF=8, pts_per_frame=250, seed=42), "config2")
            run_tracker(synth.make_track_pair())
    except Exception:  # noqa: BLE001
        say("EXCEPTION:\n" + traceback.format_exc())
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        f.write("\n".join(LINES) + "\n")


if __name__ == "__main__":
    main()
