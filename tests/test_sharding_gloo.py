"""world_size-2 gloo test (CPU) of the multi-GPU host logic of bench.py / the C ABI slicer:
rank r takes slice r of the counter-generated stream, the slices tile the global batch
exactly, and the timing reduction is a MAX over ranks.  The path has no data collective
(SURVEY 8(e)): the only communication is the barrier and the max-reduce of a scalar."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, global_batch, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import jacobi_pd_b200 as jpd
    first, count = jpd.slice_of(global_batch, world, rank)
    mine = ob.fill_synthetic(n, count, 0, 1, 4321, b0=first)                 # this rank's slice only
    evals = ob.oracle_diagonalize(mine, 1)[0]                                # CPU stand-in for the kernel
    # "time" = rank-dependent number; the bench takes the max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # gather only for the test's own verification (the product path never gathers)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([first, count]))
    chk = torch.tensor([float(evals.sum()), float(mine.sum())], dtype=torch.float64)
    allchk = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allchk, chk)
    if rank == 0:
        ret["tmax"] = float(t.item())
        ret["slices"] = [tuple(int(v) for v in s) for s in sizes]
        ret["sums"] = [tuple(float(v) for v in c) for c in allchk]
    dist.destroy_process_group()


def test_two_rank_slicing_and_max_reduce():
    world, n, global_batch = 2, 4, 2001
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, global_batch, ret), nprocs=world, join=True)
    assert ret["tmax"] == 11.0                                               # max over ranks
    (f0, c0), (f1, c1) = ret["slices"]
    assert f0 == 0 and f1 == c0 and c0 + c1 == global_batch                  # contiguous tiling
    whole = ob.fill_synthetic(n, global_batch, 0, 1, 4321)
    ev = ob.oracle_diagonalize(whole, 1)[0]
    assert np.isclose(ret["sums"][0][1] + ret["sums"][1][1], whole.sum())
    assert np.isclose(ret["sums"][0][0] + ret["sums"][1][0], ev.sum())
