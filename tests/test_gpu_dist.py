"""GPU (needs >= 2 devices): row-partitioned SpMV and CG over NVLink peer memory against the
oracle on the global system.  One process per GPU; torch.distributed (gloo) only moves the halo
index maps and the CUDA IPC handles."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, w, q):
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
    ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
    m = len(ro) - 1
    A = D.RowPartitionedCsr(ro, ci, va, rank, world, gather)
    r0, r1 = A.plan["r0"], A.plan["r1"]
    bounds_ok = np.array_equal(A.bounds, D.partition_rows(orc.merge_partition(ro, world), m))

    x = np.cos(np.arange(m) * 0.37)
    y = A.spmv(torch.from_numpy(x[r0:r1].copy()).cuda()).cpu().numpy()
    y_ref = orc.spmv_gold(ro, ci, va, x)[r0:r1]
    spmv_err = float(np.abs(y - y_ref).max() / np.abs(y_ref).max())
    dist.barrier()

    b = S.gen_rhs_rand(42, m)
    res = []
    for tol in (1e-5, 1e-9):
        it, xs, rel = A.cg_solve_single(torch.from_numpy(b[r0:r1].copy()).cuda(), 10000, tol)
        it_ref, x_ref = orc.cg_single(ro, ci, va, b, 10000, tol)
        err = float(np.abs(xs.cpu().numpy() - x_ref[r0:r1]).max() / np.abs(x_ref).max())
        res.append((it, it_ref, err, rel))
        dist.barrier()
    q.put((rank, bounds_ok, spmv_err, res))
    A.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,fused", [(2, 1), (2, 0), (4, 1), (8, 1)])
def test_row_partitioned_spmv_and_cg(gpu, world, fused, monkeypatch):
    """fused = 1: K3 stores the boundary rows of the new p straight into the neighbours' halo tails;
    fused = 0: the separate halo push kernel (the path matrices with scattered send lists take)."""
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    monkeypatch.setenv("SMLE_DIST_FUSED_PUSH", str(fused))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29700 + 10 * world + fused, 40, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bounds_ok, spmv_err, cg in res:
        assert bounds_ok, "partition rows differ from the reference merge-path search"
        assert spmv_err <= 1e-12
        for it, it_ref, err, rel in cg:
            assert abs(it - it_ref) <= max(1, round(0.02 * it_ref)), (it, it_ref)
            assert err <= 1e-6
    # every rank must report the same iteration count
    assert len({tuple(c[0] for c in r[3]) for r in res}) == 1
