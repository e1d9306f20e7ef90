"""world_size-2 gloo test of the N>1 path's host logic: shard bounds, shard-independent inputs,
counter all-reduce.  The per-shard decode is done by the oracle here (no GPU in this container);
on the GPU box the same logic drives libldpcb200 (bench.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    oracle = entry.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, _, mi = pkg.codes.config_matrix("C3")
    lo, hi = pkg.sharding.shard_range(B, rank, world)
    _, syn = oracle.sample(H, 0.05, 12345, lo, hi - lo)
    r = oracle.batch_decode(H, 0.05, mi, syn)
    ctr = torch.tensor([hi - lo, int(r["converged"].sum()), int(r["iters"].sum()), 0], dtype=torch.int64)
    pkg.sharding.allreduce_counters(ctr)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, errors=r["errors"], ctr=ctr.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path, pkg, oracle):
    B, world = 1000, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    H, _, mi = pkg.codes.config_matrix("C3")
    _, syn = oracle.sample(H, 0.05, 12345, 0, B)
    ref = oracle.batch_decode(H, 0.05, mi, syn)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert int(parts[0]["lo"]) == 0 and int(parts[0]["hi"]) == int(parts[1]["lo"]) and int(parts[1]["hi"]) == B
    assert int(parts[0]["hi"]) % 32 == 0
    assert np.array_equal(np.hstack([p["errors"] for p in parts]), ref["errors"])
    want = [B, int(ref["converged"].sum()), int(ref["iters"].sum()), 0]
    for p in parts:
        assert p["ctr"].tolist() == want


def test_shard_bounds_properties(pkg):
    for B in (0, 1, 31, 32, 33, 1000, 10_000_000):
        for w in (1, 2, 3, 4, 8):
            lo = pkg.sharding.shard_bounds(B, w)
            assert lo[0] == 0 and lo[-1] == B and all(a <= b for a, b in zip(lo, lo[1:]))
            assert all(x % 32 == 0 for x in lo[:-1])
