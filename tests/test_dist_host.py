"""CPU, world_size 2 over gloo: the host-side plumbing of the row-partitioned path
(smle_b200.dist): partition rows from merge-path coordinates, local [own | halo] systems, halo
index maps exchanged with torch.distributed, and the push plan -- checked by simulating the
pushes in numpy and comparing the assembled SpMV with the oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
    import torch.distributed as dist
    from oracle import oracle as O
    from smle_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = O.port()
    if kind == "poisson":
        ro, ci, va = orc.gen_grid3d(9, True, 6.0, -1.0)
    else:
        ro, ci, va = orc.gen_wheel(301)
        va = (np.arange(len(ci)) % 7 + 1.0)
    m = len(ro) - 1

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # the partition is the reference's merge-path search on the share diagonals
    bounds = D.partition_rows(orc.merge_partition(ro, world), m)
    plan = D.make_plan(ro, ci, va, bounds, rank, gather)
    n_local, n_halo = plan["r1"] - plan["r0"], len(plan["halo_cols"])
    x = np.cos(np.arange(m) * 0.37)
    # simulate the halo pushes: every rank publishes (dst_rank, dst_offset, values)
    pushes = []
    for peer in range(world):
        lo, hi = plan["send_off"][peer], plan["send_off"][peer + 1]
        if hi > lo:
            pushes.append((peer, int(plan["send_dst"][peer]), x[plan["r0"] + plan["send_idx"][lo:hi]]))
    x_ext = np.full(n_local + n_halo, np.nan)
    x_ext[:n_local] = x[plan["r0"]:plan["r1"]]
    for src in gather(pushes):
        for dst, off, vals in src:
            if dst == rank:
                x_ext[off:off + len(vals)] = vals
    assert not np.isnan(x_ext).any(), "a halo entry was never pushed"
    assert np.array_equal(x_ext[n_local:], x[plan["halo_cols"]]), "halo landed in the wrong order"
    y_local = orc.spmv_gold(plan["lro"], plan["lci"], plan["lva"], x_ext, n=n_local + n_halo)
    y_ref = orc.spmv_gold(ro, ci, va, x)[plan["r0"]:plan["r1"]]
    ok = bool(np.allclose(y_local, y_ref, rtol=1e-14, atol=0))
    needs = [int(v) for v in plan["needs_from"]]
    q.put((rank, ok, bounds.tolist(), n_halo, needs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["poisson", "wheel"])
def test_row_partition_plumbing_world2_gloo(kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (0 if kind == "poisson" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] and res[0][2][0] == 0          # same boundaries on both ranks
    if kind == "poisson":
        assert res[0][3] == 81 and res[1][3] == 81                  # one 9x9 plane from the neighbour
        assert res[0][4] == [0, 1] and res[1][4] == [1, 0]


def test_partition_rows_follow_reference_search():
    sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))
    from oracle import oracle as O
    from smle_b200 import dist as D
    orc = O.port()
    ro, ci, va = orc.gen_grid3d(24, True)
    coords = orc.merge_partition(ro, 8)
    rows = D.partition_rows(coords, len(ro) - 1)
    # SURVEY.md section 4 known answer for InitGrid3d(24,true), T=8
    assert rows.tolist() == [0, 1785, 3495, 5204, 6912, 8619, 10328, 12038, 13824]
