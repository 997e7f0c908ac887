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
