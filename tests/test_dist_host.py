"""CPU: the host-side plumbing of the row-partitioned path.

The product planner is C (csrc/smle_plan.cpp behind smle_dist_plan_*): a rank hands in its own rows
and gets the local [own | pad | halo] system, the halo index map and the push plan.  Here it is
checked (a) array for array against the numpy restatement in smle_b200.dist, (b) by simulating the
pushes and comparing the assembled SpMV with the oracle -- under a world_size-2 gloo group (the
blobs travel through torch.distributed) and, in one process, for 1..8 ranks on stencil, wheel and
random matrices.  No GPU: the planner is host code and the bounds come from the oracle's
MergePathSearch."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "sparse-matrix-linear-equations_b200" / "python"))


def _check_rank(D, orc, ro, ci, va, bounds, rank, world, plan_c, plan_np, x, pushes_all):
    """shared by both tests: compare the two planners and assemble this rank's SpMV from the pushes"""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    assert plan_c.n_local == r1 - r0 and plan_c.n_halo == len(plan_np["halo_cols"])
    assert plan_c.halo_base == D.halo_base(r1 - r0) and plan_c.halo_base % 16 == 0
    assert np.array_equal(plan_c.local_columns(), plan_np["lci"])
    assert np.array_equal(plan_c.halo_columns(), plan_np["halo_cols"])
    so, si, sd, nf = plan_c.send()
    assert np.array_equal(so, plan_np["send_off"]) and np.array_equal(si, plan_np["send_idx"])
    live = np.diff(so) > 0
    assert np.array_equal(sd[live], plan_np["send_dst"][live]) and np.array_equal(nf, plan_np["needs_from"])
    x_ext = np.full(plan_c.halo_base + plan_c.n_halo, np.nan)
    x_ext[:r1 - r0] = x[r0:r1]
    for src in pushes_all:
        for dst, off, vals in src:
            if dst == rank:
                x_ext[off:off + len(vals)] = vals
    halo = x_ext[plan_c.halo_base:]
    assert not np.isnan(halo).any(), "a halo entry was never pushed"
    assert np.array_equal(halo, x[plan_np["halo_cols"]]), "halo landed in the wrong order"
    x_ext = np.nan_to_num(x_ext)   # the pad between own rows and halo is never read
    y_local = orc.spmv_gold(plan_np["lro"], plan_c.local_columns(), plan_np["lva"], x_ext, n=len(x_ext))
    y_ref = orc.spmv_gold(ro, ci, va, x)[r0:r1]
    assert np.allclose(y_local, y_ref, rtol=1e-14, atol=0)


def _pushes(plan_c, x, r0):
    so, si, sd, _ = plan_c.send()
    return [(q, int(sd[q]), x[r0 + si[so[q]:so[q + 1]]]) for q in range(plan_c.world) if so[q + 1] > so[q]]


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
    plan_np = D.make_plan(ro, ci, va, bounds, rank, gather)
    r0, r1 = plan_np["r0"], plan_np["r1"]
    lo, hi = int(ro[r0]), int(ro[r1])
    plan_c = D.Plan(bounds, rank, world, m, ro[r0:r1 + 1] - lo, ci[lo:hi], gather)   # own rows only
    x = np.cos(np.arange(m) * 0.37)
    pushes_all = gather(_pushes(plan_c, x, r0))
    _check_rank(D, orc, ro, ci, va, bounds, rank, world, plan_c, plan_np, x, pushes_all)
    q.put((rank, True, [int(b) for b in bounds], plan_c.n_halo, [int(v) for v in plan_c.send()[3]]))
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


def _random_csr(rng, m, density):
    import scipy.sparse as sp
    A = sp.random(m, m, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    A = A + sp.eye(m, format="csr")
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("kind", ["poisson", "wheel", "random", "empty_rows"])
def test_c_planner_matches_numpy_planner(orc, kind, world):
    """all ranks in one process: the blobs are exchanged by hand (Plan.request_blob / Plan.finish)"""
    from smle_b200 import dist as D
    rng = np.random.default_rng(7)
    if kind == "poisson":
        ro, ci, va = orc.gen_grid3d(8, True, 6.0, -1.0)
    elif kind == "wheel":
        ro, ci, va = orc.gen_wheel(257)
    elif kind == "random":
        ro, ci, va = _random_csr(rng, 300, 0.03)
    else:
        ro, ci, va = _random_csr(rng, 200, 0.02)
        keep = np.ones(200, bool); keep[5:40] = False; keep[150:] = False      # runs of empty rows
        cnt = np.diff(ro) * keep
        sel = np.repeat(keep, np.diff(ro))
        ro, ci, va = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32), ci[sel], va[sel]
    m = len(ro) - 1
    bounds = D.partition_rows(orc.merge_partition(ro, world), m)
    x = np.cos(np.arange(m) * 0.37)
    # numpy planner, all ranks (its only communication is the gather of the request dictionaries)
    locals_ = [D.build_local_system(ro, ci, va, int(bounds[r]), int(bounds[r + 1])) for r in range(world)]
    infos = []
    for r in range(world):
        need, recv_off = D.halo_requests(locals_[r][3], np.asarray(bounds), r)
        infos.append({"n_local": int(bounds[r + 1] - bounds[r]), "need": need, "recv_off": recv_off})
    plans_np = []
    for r in range(world):
        so, si, sd, nf = D.send_plan(infos, r, int(bounds[r]))
        lro, lci, lva, hc = locals_[r]
        plans_np.append({"r0": int(bounds[r]), "r1": int(bounds[r + 1]), "lro": lro, "lci": lci, "lva": lva,
                         "halo_cols": hc, "send_off": so, "send_idx": si, "send_dst": sd, "needs_from": nf})
    # C planner: every rank sees only its rows
    plans_c = []
    for r in range(world):
        r0, r1 = int(bounds[r]), int(bounds[r + 1])
        lo, hi = int(ro[r0]), int(ro[r1])
        plans_c.append(D.Plan(bounds, r, world, m, ro[r0:r1 + 1] - lo, ci[lo:hi]))
    blobs = [p.request_blob() for p in plans_c]
    for p in plans_c:
        p.finish(blobs)
    pushes_all = [_pushes(p, x, int(bounds[r])) for r, p in enumerate(plans_c)]
    for r in range(world):
        _check_rank(D, orc, ro, ci, va, bounds, r, world, plans_c[r], plans_np[r], x, pushes_all)
    for p in plans_c:
        p.close()


def test_planner_rejects_bad_input(orc):
    from smle_b200 import SmleError
    from smle_b200 import dist as D
    ro, ci, va = orc.gen_grid3d(4, True)
    m = len(ro) - 1
    with pytest.raises(SmleError):
        D.Plan([0, m // 2, m - 1], 0, 2, m, ro[:m // 2 + 1], ci[:ro[m // 2]])          # bounds do not cover the columns
    with pytest.raises(SmleError):
        D.Plan([0, m // 2, m], 0, 2, m, ro[:m // 2 + 1], ci[:ro[m // 2]] + m)          # column out of range
    p = D.Plan([0, m // 2, m], 0, 2, m, ro[:m // 2 + 1], ci[:ro[m // 2]])
    with pytest.raises(SmleError):
        p.send()                                                                        # not finished yet
    p.close()


def test_partition_rows_follow_reference_search():
    from oracle import oracle as O
    from smle_b200 import dist as D
    orc = O.port()
    ro, ci, va = orc.gen_grid3d(24, True)
    coords = orc.merge_partition(ro, 8)
    rows = D.partition_rows(coords, len(ro) - 1)
    # SURVEY.md section 4 known answer for InitGrid3d(24,true), T=8
    assert rows.tolist() == [0, 1785, 3495, 5204, 6912, 8619, 10328, 12038, 13824]


def test_slab_generator_matches_global_generator(S, orc):
    """rows [r0, r1) from smle_gen_grid3d_rows == the same rows of the global generator (itself pinned
    to the reference's InitGrid3d + CsrMatrix::Init in test_capi_host.py); the stream range likewise."""
    w = 11
    ro, ci, va = S.gen_grid3d(w, True, 6.0, -1.0)
    assert np.array_equal(S.gen_grid3d_row_offsets(w, True), ro)
    m = len(ro) - 1
    for r0, r1 in ((0, m), (0, 1), (17, 17), (100, 731), (m - 5, m)):
        lro, lci, lva = S.gen_grid3d_rows(w, r0, r1, int(ro[r1] - ro[r0]), True, 6.0, -1.0)
        assert np.array_equal(lro, ro[r0:r1 + 1] - ro[r0])
        assert np.array_equal(lci, ci[ro[r0]:ro[r1]]) and np.array_equal(lva, va[ro[r0]:ro[r1]])
    b = S.gen_rhs_rand(42, 1000)
    assert np.array_equal(S.gen_rhs_rand_range(42, 123, 456), b[123:579])
