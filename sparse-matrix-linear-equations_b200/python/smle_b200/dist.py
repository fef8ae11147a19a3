"""Row-partitioned SpMV / CG across GPUs (one process per GPU) -- host-side plumbing.

Net-new relative to the reference (no distributed code there); SURVEY.md section 8(e):

* the row partition is cut at the reference's merge-path coordinates: part g owns rows
  [x_g, x_{g+1}) with x_g = MergePathSearch(min(g*ceil((m+nnz)/G), m+nnz)).x -- a row cut mid-way
  belongs whole to the later part;
* the local matrix keeps its rows and remaps columns to [own | halo], halo = sorted unique
  out-of-range columns (for 3-D Poisson slabs: one w*w plane per neighbour);
* the halo index maps are exchanged once through torch.distributed (gloo or nccl -- plumbing
  only); the data path is the peer-memory kernels of csrc/smle_dist.cuh.

Everything in the first half of this file is pure numpy and is exercised on CPU with a
world_size-2 gloo group (tests/test_dist_host.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_I, _P, _D = C.c_int, C.c_void_p, C.c_double


# ---------------------------------------------------------------------------------------------
# host logic (numpy only)
# ---------------------------------------------------------------------------------------------
def partition_rows(coords: np.ndarray, num_rows: int) -> np.ndarray:
    """share boundaries (G+1, 2) of the merge path -> first row of every part, shape (G+1,)."""
    rows = np.asarray(coords)[:, 0].astype(np.int64).copy()
    rows[0], rows[-1] = 0, num_rows
    return rows


def build_local_system(ro, ci, va, r0: int, r1: int):
    """rows [r0, r1) of the global CSR with columns remapped to [own | halo].

    returns (local_row_offsets, local_column_indices, local_values, halo_cols) where halo_cols are
    the sorted unique global columns outside [r0, r1) this part touches."""
    lo, hi = int(ro[r0]), int(ro[r1])
    lro = (np.asarray(ro[r0:r1 + 1], dtype=np.int64) - lo).astype(np.int32)
    cols = np.asarray(ci[lo:hi], dtype=np.int64)
    own = (cols >= r0) & (cols < r1)
    halo_cols = np.unique(cols[~own])
    n_local = r1 - r0
    lci = np.where(own, cols - r0, n_local + np.searchsorted(halo_cols, cols)).astype(np.int32)
    return lro, lci, np.ascontiguousarray(va[lo:hi]), halo_cols.astype(np.int64)


def halo_requests(halo_cols: np.ndarray, bounds: np.ndarray, rank: int):
    """split this rank's halo columns by owner.  returns (need, recv_off): need[q] = global columns
    wanted from rank q (sorted), recv_off[q] = position of that group inside the halo region."""
    world = len(bounds) - 1
    owner = np.searchsorted(bounds, halo_cols, side="right") - 1
    need, recv_off = {}, {}
    for q in range(world):
        sel = halo_cols[owner == q]
        need[q] = sel
        recv_off[q] = int(np.searchsorted(halo_cols, sel[0])) if len(sel) else 0
    assert len(need[rank]) == 0
    return need, recv_off


def send_plan(infos, rank: int, r0: int):
    """From every rank's published {n_local, need, recv_off}: what THIS rank pushes where.
    returns (send_off[world+1], send_idx, send_dst[world], needs_from[world])."""
    world = len(infos)
    send_off, idx, send_dst, needs_from = [0], [], [], []
    for q in range(world):
        want = np.asarray(infos[q]["need"].get(rank, np.zeros(0, np.int64)), dtype=np.int64) if q != rank else np.zeros(0, np.int64)
        idx.append((want - r0).astype(np.int32))
        send_off.append(send_off[-1] + len(want))
        send_dst.append(int(infos[q]["n_local"] + infos[q]["recv_off"].get(rank, 0)))
        needs_from.append(1 if q != rank and len(infos[rank]["need"].get(q, ())) > 0 else 0)
    send_idx = np.concatenate(idx) if idx else np.zeros(0, np.int32)
    return (np.asarray(send_off, np.int32), np.ascontiguousarray(send_idx, np.int32),
            np.asarray(send_dst, np.int32), np.asarray(needs_from, np.int32))


def make_plan(ro, ci, va, bounds, rank, all_gather_object):
    """Everything a rank needs: local system + push plan.  `all_gather_object(obj) -> list` is the
    only communication (torch.distributed.all_gather_object wrapped by the caller)."""
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    lro, lci, lva, halo_cols = build_local_system(ro, ci, va, r0, r1)
    need, recv_off = halo_requests(halo_cols, np.asarray(bounds), rank)
    infos = all_gather_object({"n_local": r1 - r0, "need": need, "recv_off": recv_off})
    send_off, send_idx, send_dst, needs_from = send_plan(infos, rank, r0)
    return {"r0": r0, "r1": r1, "lro": lro, "lci": lci, "lva": lva, "halo_cols": halo_cols,
            "send_off": send_off, "send_idx": send_idx, "send_dst": send_dst, "needs_from": needs_from}


# ---------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------
class RowPartitionedCsr:
    """This rank's block of a row-partitioned CsrMatrix plus the NVLink peer connections."""

    def __init__(self, ro, ci, va, rank: int, world: int, all_gather_object, bounds=None):
        ro = np.ascontiguousarray(ro, dtype=np.int32)
        m = len(ro) - 1
        if bounds is None:
            # merge-path coordinates on the GPU (bit-exact with the reference's MergePathSearch)
            bounds = partition_rows(capi.merge_path_partition(ro, world), m)
        self.bounds = np.asarray(bounds)
        self.rank, self.world, self.num_rows_global = rank, world, m
        p = make_plan(ro, ci, va, self.bounds, rank, all_gather_object)
        self.plan = p
        self.n_local = p["r1"] - p["r0"]
        self.n_halo = len(p["halo_cols"])
        self.local = capi.CsrMatrix(p["lro"], p["lci"], p["lva"], num_cols=self.n_local + self.n_halo)
        L = capi.lib()
        h = _P()
        capi._check(L.smle_dist_create(C.byref(h), self.local._h, _I(rank), _I(world), _I(self.n_local), _I(self.n_halo),
                                       p["send_off"].ctypes.data_as(_P), p["send_idx"].ctypes.data_as(_P),
                                       p["send_dst"].ctypes.data_as(_P), p["needs_from"].ctypes.data_as(_P)))
        self._h = h
        mine = (C.c_ubyte * 64)()
        capi._check(L.smle_dist_ipc_handle(self._h, mine))
        handles = all_gather_object(bytes(mine))
        blob = b"".join(handles)
        capi._check(L.smle_dist_connect(self._h, blob))

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().smle_dist_destroy(self._h)
            self._h = None
            self.local.close()

    def spmv(self, x_local, out=None):
        """y_local = (A x)_local; collective (call it on every rank, barrier between calls)."""
        import torch
        y = out if out is not None else torch.empty(self.n_local, dtype=torch.float64, device=x_local.device)
        px, _, _ = capi._arg(x_local, np.float64)
        py, _, _ = capi._arg(y, np.float64, writable=True)
        capi._check(capi.lib().smle_dist_spmv_f64(self._h, px, py))
        capi.sync()
        return y

    def cg_solve_single(self, b_local, max_iters: int, tolerance: float, out=None):
        """-> (iterations, x_local, final_rel_res); CGSolveSingle semantics on the global system."""
        import torch
        x = out if out is not None else torch.empty(self.n_local, dtype=torch.float64, device=b_local.device)
        pb, _, _ = capi._arg(b_local, np.float64)
        px, _, _ = capi._arg(x, np.float64, writable=True)
        it, rel = _I(0), _D(0)
        capi._check(capi.lib().smle_dist_cg_f64(self._h, pb, px, _I(max_iters), _D(tolerance), C.byref(it), C.byref(rel)))
        return it.value, x, rel.value
