"""CPU, world_size 2 over gloo: host logic of the column-sharded multi-RHS CG (first row).
Every rank runs the ORACLE's CGSolveMultiple on its columns; the merged (iterations, history) must equal
what the oracle returns for the whole block -- the statement the GPU path relies on when it shards
the k right-hand sides over ranks with one gather at the end."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
    import torch.distributed as dist
    from oracle import oracle as O
    from smle_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = O.port()
    ro, ci, va = orc.gen_grid3d(10, True, 6.0, -1.0)
    n, k = len(ro) - 1, 6
    B = orc.rhs_rand(42, n * k).reshape(n, k)
    B[:, 1] *= 1e-3          # columns converge at different iterations
    lo, hi = D.shard_columns(k, rank, world)
    it, X, hist = orc.cg_multi(ro, ci, va, np.ascontiguousarray(B[:, lo:hi]), hi - lo, 10000, 1e-7, O.MERGE, 4)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    iters, hist_all = D.merge_sharded_results(gather((it, hist)))
    it_full, X_full, hist_full = orc.cg_multi(ro, ci, va, B, k, 10000, 1e-7, O.MERGE, 4)
    ok = iters == it_full and len(hist_all) == len(hist_full) and np.allclose(hist_all, hist_full, rtol=1e-9) \
        and np.allclose(X, X_full[:, lo:hi], rtol=1e-9, atol=1e-14)
    q.put((rank, bool(ok), iters, it_full, it))
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharded_merge_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29660, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] == res[0][3]


def test_shard_columns_cover_the_block():
    from smle_b200 import dist as D
    for k in (1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            cuts = [D.shard_columns(k, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == k
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def test_merge_keeps_frozen_residuals():
    from smle_b200 import dist as D
    it, h = D.merge_sharded_results([(3, [0.5, 0.2, 0.09]), (5, [0.4, 0.3, 0.2, 0.1, 0.05])])
    assert it == 5 and np.allclose(h, [0.5, 0.3, 0.2, 0.1, 0.09])
