"""world_size-2 gloo test (CPU) of the N > 1 path's host logic: bench.py's shard partition, the max-over-ranks
timing reduction and the 'no data-path collective' property (each rank's shard is processed independently and the
union of shards is the whole batch).  The per-shard compute stands in with the CPU oracle here; on the GPU box the
same partition feeds gschur_cuda_batched per rank."""
import os
import sys

import numpy as np

from common import ROOT


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total, n = 13, 8
    rng = np.random.default_rng(42)
    A = np.asfortranarray(rng.random((n, n, total)))
    lo, hi = bench.shard_bounds(total, world, rank)
    T, Z, w, info = O.gschur_batched(np.asfortranarray(A[:, :, lo:hi].copy(order="F")), 0, nthreads=1)
    # gather (host-side, outside any timed region): shard sizes + eigenvalues
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    pad = torch.zeros((n, total), dtype=torch.complex128)
    pad[:, : hi - lo] = torch.from_numpy(w)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)     # fake per-rank elapsed ms
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        got = np.concatenate([out[r][:, : int(sizes[r])].numpy() for r in range(world)], axis=1)
        _, _, wfull, info_full = O.gschur_batched(A.copy(order="F"), 0, nthreads=1)
        np.save(os.path.join(tmpdir, "ok.npy"),
                np.array([float(np.array_equal(got, wfull)), float(sum(int(s) for s in sizes) == total), t.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_batch():
    sys.path.insert(0, ROOT)
    import bench
    for total in (0, 1, 7, 65536):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = bench.shard_bounds(total, world, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == total


def test_two_rank_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok[0] == 1.0 and ok[1] == 1.0
    assert ok[2] == 11.0       # max over ranks
