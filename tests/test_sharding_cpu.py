"""N>1 host logic on CPU (gloo, world_size 2): point sharding partitions the window, and the sum over ranks of the
per-shard reduced systems equals the full window's system — the property the one NCCL all-reduce per GN step relies on
(SURVEY.md §8e). The per-shard systems come from the CPU oracle here; the GPU version of this test is in
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ldso_b200 import synth
from tests.parity import rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_window_partitions_points():
    win = synth.make_window(nF=4, pts_per_frame=30, w=320, h=240, seed=9)
    for world in (2, 3, 8):
        shards = [synth.shard_window(win, r, world) for r in range(world)]
        assert sum(s.nP for s in shards) == win.nP
        assert sum(s.nR for s in shards) == win.nR
        us = np.concatenate([np.stack([s.pt_host, s.pt_u, s.pt_v], 1) for s in shards])
        full = np.stack([win.pt_host, win.pt_u, win.pt_v], 1)
        assert sorted(map(tuple, us.tolist())) == sorted(map(tuple, full.tolist()))
        for s in shards:
            assert np.all(np.diff(s.pt_host) >= 0)          # still ordered by host
            assert s.res_begin[-1] == s.nR
        # newest-frame slot bookkeeping used by ldso_b200_set_shard
        counts = [int(np.sum(s.res_target == win.nF - 1)) for s in shards]
        assert sum(counts) == int(np.sum(win.res_target == win.nF - 1))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_py
    win = synth.make_window(nF=4, pts_per_frame=40, w=320, h=240, seed=9)
    sh = synth.shard_window(win, rank, world)
    o = oracle_py.OracleBA(sh, threads_mode=1)
    e = o.optimize_begin()
    o.solve_system(0)
    s = o.system()
    n = s["HA"].shape[0]
    buf = torch.from_numpy(np.concatenate([s["HA"].ravel(), s["bA"], s["Hsc"].ravel(), s["bsc"], [e]]))
    dist.all_reduce(buf)      # the one collective per GN step
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_shards_equals_full_window():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    red = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from tests import oracle_py
    win = synth.make_window(nF=4, pts_per_frame=40, w=320, h=240, seed=9)
    o = oracle_py.OracleBA(win, threads_mode=1)
    e = o.optimize_begin()
    o.solve_system(0)
    s = o.system()
    n = s["HA"].shape[0]
    k = 0
    HA = red[k:k + n * n].reshape(n, n); k += n * n
    bA = red[k:k + n]; k += n
    Hsc = red[k:k + n * n].reshape(n, n); k += n * n
    bsc = red[k:k + n]; k += n
    # column-major ravel of symmetric-ish matrices: compare against the same ravel of the full window
    assert rel_err(HA.ravel(), s["HA"].ravel()) < 1e-6
    assert rel_err(bA, s["bA"]) < 1e-6
    assert rel_err(Hsc.ravel(), s["Hsc"].ravel()) < 1e-6
    assert rel_err(bsc, s["bsc"]) < 1e-6
    assert abs(red[k] - e) < 1e-6 * abs(e)


def test_sequence_generator_config4_inputs():
    """ldso_b200.seq (BASELINE configs[3]'s synthetic sequence, host side only): sliding-window bookkeeping, the flattened window's CSR
    layout, and geometric consistency of poses, depths and rendered images (a keyframe's point at its true depth reprojects into the
    next keyframe onto the same scene intensity, up to the frames' affine brightness)."""
    from ldso_b200 import seq as seqmod
    K = seqmod.KITTI_K * np.array([0.25, 0.25, 0.25, 0.25])
    s = seqmod.make_sequence(n_frames=11, w=308, h=92, K=K, kf_every=5, window=2, pts_per_kf=40, seed=3, outlier_frac=0.0)
    assert s.levels == synth.pyr_levels_for(308, 92) and len(s.images) == 11 and sorted(s.kf_points) == [0, 5, 10]
    assert s.window_kfs(0) == [0] and s.window_kfs(5) == [0, 5] and s.window_kfs(10) == [5, 10]
    R, t = s.rel_pose(5, 10)
    assert np.allclose(R @ s.Rcw[5], s.Rcw[10]) and np.allclose(R @ s.tcw[5] + t, s.tcw[10])
    W = seqmod.window_arrays(s, [5, 10])
    nP = len(W["pt_host"])
    assert nP == 80 and W["res_begin"][0] == 0 and W["res_begin"][-1] == len(W["res_target"]) == nP
    assert np.all(np.diff(W["res_begin"]) == 1) and np.all(W["res_target"] != W["pt_host"])
    assert list(W["frame_id"]) == [1, 2] and W["state"].shape == (2, 10) and np.all(W["pt_idepth"] > 0)
    # a point of keyframe 5 at its (noisy, 1 %) inverse depth, reprojected into keyframe 10
    P = s.kf_points[5]
    fx, fy, cx, cy = K
    x = (P["u"] - cx) / fx; y = (P["v"] - cy) / fy
    X = np.stack([x, y, np.ones_like(x)], 1) / P["idepth_zero"][:, None]
    Y = X @ R.T + t
    u2 = fx * Y[:, 0] / Y[:, 2] + cx; v2 = fy * Y[:, 1] / Y[:, 2] + cy
    ok = (u2 > 2) & (u2 < s.w - 3) & (v2 > 2) & (v2 < s.h - 3)
    assert ok.sum() >= 20
    I5 = synth.sample_bilin(synth.make_images(s.images[5], 1)[0], P["u"][ok].astype(np.float64), P["v"][ok].astype(np.float64))[0]
    I10 = synth.sample_bilin(synth.make_images(s.images[10], 1)[0], u2[ok], v2[ok])[0]
    a5, b5 = s.aff[5]; a10, b10 = s.aff[10]
    scene5 = (I5 - b5) / np.exp(a5); scene10 = (I10 - b10) / np.exp(a10)
    assert np.median(np.abs(scene5 - scene10)) < 3.0        # grey levels; 1 % depth noise moves the reprojection by a fraction of a pixel
