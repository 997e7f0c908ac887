"""Sharded (multi-GPU) GN iteration vs the single-context result, run under torchrun (one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py
Every rank holds a point shard; one NCCL all-reduce of the reduced system per GN step. Rank 0 also runs the
full window on its own GPU in a second context and compares. Exit code != 0 on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldso_b200 import capi, synth  # noqa: E402
from tests.parity import rel_err  # noqa: E402


class _DevBuf:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def config3_vs_oracle(rank, world, lr):
    """BASELINE configs[2]: the 8 KF x 20 000-point window sharded over the job's GPUs (device-side peer exchange), first GN iteration
    against the CPU ORACLE on rank 0: energy, lastHS, lastbS to 1e-4 (north_star), the update off the gauge direction."""
    full = synth.make_window(nF=8, pts_per_frame=2500, seed=42)
    win = synth.shard_window(full, rank, world)
    ctx = capi.Context(win.w, win.h, win.levels, device=lr)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.load_synth_window(win)
    counts = [int(np.sum(synth.shard_window(full, r, world).res_target == full.nF - 1)) for r in range(world)]
    ctx.set_shard(int(np.sum(counts[:rank])), int(np.sum(counts)))
    handles = [None] * world
    dist.all_gather_object(handles, ctx.peer_export())
    ctx.peer_connect(rank, world, handles)
    dist.barrier()
    e0 = ctx.optimize_begin()
    ctx.gn_iterations(0, 1)
    sol = ctx.last_solution()
    e1 = ctx.energy()[0]
    ok = ctx.peer_error() == 0
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        from tests import oracle_py
        o = oracle_py.OracleBA(full, threads_mode=0)
        eo0 = o.optimize_begin()
        o.gn_iteration(0)
        so = o.system()
        P = o.nullspace_projector(); I = np.eye(P.shape[0])
        eh, eb = rel_err(sol["lastHS"], so["lastHS"]), rel_err(sol["lastbS"], so["lastbS"])
        ex = rel_err((I - P) @ sol["lastX"], (I - P) @ so["lastX"])
        eo1 = o.L.oracle_ba_last_energy(o.o)
        print(f"config3 x{world} vs oracle: energy0 {abs(e0 - eo0) / abs(eo0):.2e} lastHS {eh:.2e} lastbS {eb:.2e} lastX(gauge-proj) {ex:.2e} energy1 {abs(e1 - eo1) / abs(eo1):.2e}")
        ok &= abs(e0 - eo0) <= 1e-5 * abs(eo0) and eh < 1e-4 and eb < 1e-4 and ex < 1e-4 and abs(e1 - eo1) <= 2e-3 * abs(eo1)
        print("MULTI_CHECK_CONFIG3", "OK" if ok else "FAIL", "world", world)
    ctx.close()
    return ok


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    if "--config3" in sys.argv:
        ok = config3_vs_oracle(rank, world, lr)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag.item()) == 1 else 1)
    full = synth.make_window(nF=6, pts_per_frame=120, w=320, h=240, seed=17)
    win = synth.shard_window(full, rank, world)
    ctx = capi.Context(win.w, win.h, win.levels, device=lr)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.load_synth_window(win)
    counts = [int(np.sum(synth.shard_window(full, r, world).res_target == full.nF - 1)) for r in range(world)]
    ctx.set_shard(int(np.sum(counts[:rank])), int(np.sum(counts)))
    ptr, n = ctx.reduce_buffer()
    red = torch.as_tensor(_DevBuf(ptr, n), device=f"cuda:{lr}")
    sols, energies = [], []
    for it in range(-1, 3):
        ctx.gn_phase_a(it)
        dist.all_reduce(red)
        ctx.gn_phase_b()
        if it >= 0:
            sols.append(ctx.last_solution())
        energies.append(ctx.energy()[0])
    pts = ctx.points()
    ok = True
    # ---- the same iterations with the device-side exchange (k2r_peer_allreduce over NVLink peer memory): no NCCL call in the loop
    ctx2 = capi.Context(win.w, win.h, win.levels, device=lr)
    ctx2.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx2.load_synth_window(win)
    ctx2.set_shard(int(np.sum(counts[:rank])), int(np.sum(counts)))
    handles = [None] * world
    dist.all_gather_object(handles, ctx2.peer_export())
    ctx2.peer_connect(rank, world, handles)
    dist.barrier()
    e2 = [ctx2.optimize_begin()]
    sols2 = []
    for it in range(3):
        ctx2.gn_iterations(it, 1)
        sols2.append(ctx2.last_solution())
        e2.append(ctx2.energy()[0])
    ok &= ctx2.peer_error() == 0
    for it in range(3):
        for k in ("lastHS", "lastbS", "lastX"):
            same = np.array_equal(sols2[it][k], sols[it][k]) if world == 2 else rel_err(sols2[it][k], sols[it][k]) < 1e-9
            if not same:
                print(f"rank {rank}: peer path differs from the NCCL path at it{it} {k}: {rel_err(sols2[it][k], sols[it][k]):.3e}")
                ok = False
    ok &= bool(np.allclose(e2, energies, rtol=1e-12))
    if rank == 0:
        print("peer-exchange path: energies", [round(x, 3) for x in e2])
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        ref = capi.Context(full.w, full.h, full.levels, device=lr)
        ref.load_synth_window(full)
        e_ref = [ref.optimize_begin()]
        P = ref.nullspace_projector()
        I = np.eye(P.shape[0])
        for it in range(3):
            ref.gn_iterations(it, 1)
            s = ref.last_solution()
            e_ref.append(ref.energy()[0])
            eh = rel_err(sols[it]["lastHS"], s["lastHS"]); eb = rel_err(sols[it]["lastbS"], s["lastbS"])
            ex = rel_err((I - P) @ sols[it]["lastX"], (I - P) @ s["lastX"])
            print(f"it{it}: lastHS {eh:.2e} lastbS {eb:.2e} lastX(gauge-proj) {ex:.2e} energy {energies[it + 1]:.3f} vs {e_ref[it + 1]:.3f}")
            if it == 0:
                ok &= eh < 1e-6 and eb < 1e-5 and ex < 1e-4
        ok &= abs(energies[0] - e_ref[0]) <= 1e-6 * abs(e_ref[0])
        ok &= np.allclose(energies, e_ref, rtol=2e-3)
        ok &= rel_err(ctx.frames()["frameEnergyTH"], ref.frames()["frameEnergyTH"]) < 1e-3
        print("MULTI_CHECK", "OK" if ok else "FAIL", "world", world)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
