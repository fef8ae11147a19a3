"""GPU (needs >= 2 devices): row-partitioned SpMV and CG over NVLink peer memory against the
oracle on the global system.  One process per GPU; torch.distributed (gloo) only moves the request
blobs of the planner and the CUDA IPC handles."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _spd_random(m, seed):
    """diagonally dominant symmetric matrix with a scattered pattern: halo columns everywhere, send
    lists that are not contiguous runs (the separate push kernel is chosen automatically)"""
    import scipy.sparse as sp
    W = sp.random(m, m, density=6.0 / m, random_state=np.random.RandomState(seed), format="csr")
    W = W + W.T
    A = (sp.diags(np.asarray(W.sum(axis=1)).ravel() + 1.0) - W).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
    import torch
    import torch.distributed as dist
    import smle_b200 as S
    from smle_b200 import dist as D
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    S.init(rank)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    orc = O.port()
    w = 40
    if kind == "random":
        ro, ci, va = _spd_random(30000, 5)
    else:
        ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
    m = len(ro) - 1
    if kind == "slab":
        A = D.RowPartitionedCsr.grid3d(w, rank, world, gather)          # this rank generates its rows only
    else:
        A = D.RowPartitionedCsr.from_global(ro, ci, va, rank, world, gather)
    r0, r1 = A.r0, A.r1
    bounds_ok = np.array_equal(A.bounds, D.partition_rows(orc.merge_partition(ro, world), m))

    x = np.cos(np.arange(m) * 0.37)
    y = A.spmv(torch.from_numpy(x[r0:r1].copy()).cuda()).cpu().numpy()
    y_ref = orc.spmv_gold(ro, ci, va, x)[r0:r1]
    spmv_err = float(np.abs(y - y_ref).max() / np.abs(y_ref).max())
    dist.barrier()

    b = S.gen_rhs_rand(42, m)
    assert np.array_equal(S.gen_rhs_rand_range(42, r0, r1 - r0), b[r0:r1])
    res = []
    for tol, host in ((1e-5, False), (1e-9, False), (1e-7, True)):
        if host:   # host buffers through the C ABI (is_device_ptr = 0)
            it, xs, rel = A.cg_solve_single(b[r0:r1].copy(), 10000, tol)
        else:
            it, xs, rel = A.cg_solve_single(torch.from_numpy(b[r0:r1].copy()).cuda(), 10000, tol)
            xs = xs.cpu().numpy()
        it_ref, x_ref = orc.cg_single(ro, ci, va, b, 10000, tol)
        err = float(np.abs(xs - x_ref[r0:r1]).max() / np.abs(x_ref).max())
        res.append((it, it_ref, err, rel))
        dist.barrier()
    # max_iters cap: every rank stops after exactly 7 iterations
    it, _, _ = A.cg_solve_single(torch.from_numpy(b[r0:r1].copy()).cuda(), 7, 1e-30)
    res.append((it, 7, 0.0, 0.0))
    dist.barrier()
    q.put((rank, bounds_ok, spmv_err, res))
    A.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,fused", [(2, "global", 1), (2, "global", 0), (2, "slab", 1), (2, "random", 1),
                                              (4, "slab", 1), (4, "random", 1), (8, "slab", 1), (8, "global", 0)])
def test_row_partitioned_spmv_and_cg(gpu, world, kind, fused, monkeypatch):
    """fused = 1: K3 stores the boundary rows of the new p straight into the neighbours' halo tails;
    fused = 0: the separate halo push kernel (the path matrices with scattered send lists take)."""
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    monkeypatch.setenv("SMLE_DIST_FUSED_PUSH", str(fused))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + 10 * world + fused + 2 * ["global", "slab", "random"].index(kind)
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=300) for _ in range(world))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    for rank, bounds_ok, spmv_err, cg in res:
        assert bounds_ok, "partition rows differ from the reference merge-path search"
        assert spmv_err <= 1e-12
        for it, it_ref, err, rel in cg:
            assert abs(it - it_ref) <= max(1, round(0.02 * it_ref)), (it, it_ref)
            assert err <= 1e-6
    # every rank must report the same iteration counts
    assert len({tuple(c[0] for c in r[3]) for r in res}) == 1


def test_world1_partition_is_the_plain_solver(gpu, orc):
    """a partition of one rank goes through the same kernels (mailbox to itself, no halo)"""
    import torch
    from smle_b200 import dist as D
    A = D.RowPartitionedCsr.grid3d(24, 0, 1, lambda obj: [obj])
    ro, ci, va = gpu.gen_grid3d(24, True, 6.0, -1.0)
    n = len(ro) - 1
    assert A.n_local == n and A.n_halo == 0
    b = gpu.gen_rhs_rand(42, n)
    it, x, rel = A.cg_solve_single(torch.from_numpy(b).cuda(), 10000, 1e-8)
    it_ref, x_ref = orc.cg_single(ro, ci, va, b, 10000, 1e-8)
    assert abs(it - it_ref) <= max(1, round(0.02 * it_ref))
    assert np.abs(x.cpu().numpy() - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    # host buffers
    it2, x2, _ = A.cg_solve_single(b, 10000, 1e-8)
    assert it2 == it and np.array_equal(x2, x.cpu().numpy())
    A.close()


def _sharded_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
    import torch
    import torch.distributed as dist
    import smle_b200 as S
    from smle_b200 import dist as D
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    S.init(rank)

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    orc = O.port()
    ro, ci, va = S.gen_grid3d(20, True, 6.0, -1.0)
    n, k = len(ro) - 1, 8
    B = S.gen_rhs_rand(42, n * k).reshape(n, k).copy()
    B[:, 5] *= 1e-4                                  # one column converges much earlier than the rest
    lo, hi = D.shard_columns(k, rank, world)
    a = S.CsrMatrix(ro, ci, va)                      # A replicated
    iters, X, hist = D.cg_solve_multiple_column_sharded(a, np.ascontiguousarray(B[:, lo:hi]), 10000, 1e-8, gather)
    it_o, X_o, hist_o = orc.cg_multi(ro, ci, va, B, k, 10000, 1e-8, O.MERGE, 8)
    nh = min(len(hist), len(hist_o)) - 2
    ok = abs(iters - it_o) <= max(1, round(0.02 * it_o)) and abs(len(hist) - len(hist_o)) <= 1 \
        and np.allclose(hist[:nh], hist_o[:nh], rtol=1e-5) and np.allclose(X, X_o[:, lo:hi], rtol=1e-6, atol=1e-9)
    q.put((rank, bool(ok), iters, it_o))
    a.close()
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharded_multi_rhs_cg(gpu):
    """SURVEY.md section 8e, first row: columns of the k right-hand sides over 2 GPUs, A replicated; iteration count
    and error history of the whole block from one gather, against CGSolveMultiple's restatement on all k columns."""
    if gpu.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, 29790, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=300) for _ in range(2))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]
