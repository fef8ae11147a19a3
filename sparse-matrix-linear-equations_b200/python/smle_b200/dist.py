"""Row-partitioned SpMV / CG across GPUs (one process per GPU) -- host-side plumbing.

Net-new relative to the reference (no distributed code there); SURVEY.md section 8(e):

* the row partition is cut at the reference's merge-path coordinates: part g owns rows
  [x_g, x_{g+1}) with x_g = MergePathSearch(min(g*ceil((m+nnz)/G), m+nnz)).x -- a row cut mid-way
  belongs whole to the later part;
* the local matrix keeps its rows and remaps columns to [own | halo], halo = sorted unique
  out-of-range columns (for 3-D Poisson slabs: one w*w plane per neighbour);
* the halo index maps are exchanged once through torch.distributed (gloo or nccl -- plumbing
  only); the data path is the peer-memory kernels of csrc/smle_dist.cuh.

The product path is the C planner behind the C ABI (csrc/smle_plan.cpp, `Plan` below): a rank hands
in its own rows only.  The numpy functions in the first half restate the same plan; the CPU tests
(tests/test_dist_host.py, world_size-2 gloo group) compare the two array for array.
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


def halo_base(n_local: int) -> int:
    """first halo column of the local system: the halo tail starts on a 128-byte line of its own"""
    return (n_local + 15) // 16 * 16


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
    lci = np.where(own, cols - r0, halo_base(n_local) + np.searchsorted(halo_cols, cols)).astype(np.int32)
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
        send_dst.append(int(halo_base(infos[q]["n_local"]) + infos[q]["recv_off"].get(rank, 0)))
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
# the C planner (csrc/smle_plan.cpp) -- what the product uses; the numpy functions above restate it
# for the CPU tests
# ---------------------------------------------------------------------------------------------
class Plan:
    """smle_dist_plan_*: local system + push plan of one rank, from ITS rows only."""

    def __init__(self, bounds, rank: int, world: int, n_global: int, lro, ci_global, all_gather_object=None):
        """all_gather_object(bytes) -> list of every rank's bytes moves the request blobs; without it the
        caller exchanges `request_blob()` itself and calls `finish(blobs)`."""
        L = capi.lib()
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int32)
        lro = np.ascontiguousarray(lro, dtype=np.int32)
        ci_global = np.ascontiguousarray(ci_global, dtype=np.int32)
        h = _P()
        capi._check(L.smle_dist_plan_create(C.byref(h), _I(rank), _I(world), self.bounds.ctypes.data_as(_P), _I(n_global),
                                            lro.ctypes.data_as(_P), ci_global.ctypes.data_as(_P)))
        self._h = h
        nl, nh, hb, nz = _I(0), _I(0), _I(0), _I(0)
        capi._check(L.smle_dist_plan_dims(h, C.byref(nl), C.byref(nh), C.byref(hb), C.byref(nz)))
        self.n_local, self.n_halo, self.halo_base, self.nnz_local = nl.value, nh.value, hb.value, nz.value
        self.rank, self.world = rank, world
        if all_gather_object is not None:
            self.finish(all_gather_object(self.request_blob()))

    def request_blob(self) -> bytes:
        L = capi.lib()
        blob = np.zeros(int(L.smle_dist_plan_request_size(self._h)), dtype=np.int32)
        capi._check(L.smle_dist_plan_request(self._h, blob.ctypes.data_as(_P)))
        return blob.tobytes()

    def finish(self, blobs) -> None:
        blobs = [np.frombuffer(b, dtype=np.int32) for b in blobs]
        off = np.zeros(self.world + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(b) for b in blobs])
        allb = np.ascontiguousarray(np.concatenate(blobs), dtype=np.int32)
        capi._check(capi.lib().smle_dist_plan_finish(self._h, allb.ctypes.data_as(_P), off.ctypes.data_as(_P)))

    def local_columns(self):
        out = np.zeros(self.nnz_local, dtype=np.int32)
        capi._check(capi.lib().smle_dist_plan_local_columns(self._h, out.ctypes.data_as(_P)))
        return out

    def halo_columns(self):
        out = np.zeros(self.n_halo, dtype=np.int32)
        capi._check(capi.lib().smle_dist_plan_halo_columns(self._h, out.ctypes.data_as(_P)))
        return out

    def send(self):
        """(send_off, send_idx, send_dst, needs_from)"""
        L = capi.lib()
        so = np.zeros(self.world + 1, dtype=np.int32)
        sd = np.zeros(self.world, dtype=np.int32)
        nf = np.zeros(self.world, dtype=np.int32)
        capi._check(L.smle_dist_plan_send(self._h, so.ctypes.data_as(_P), None, sd.ctypes.data_as(_P), nf.ctypes.data_as(_P)))
        si = np.zeros(int(so[-1]), dtype=np.int32)
        capi._check(L.smle_dist_plan_send(self._h, None, si.ctypes.data_as(_P), None, None))
        return so, si, sd, nf

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().smle_dist_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------
class RowPartitionedCsr:
    """This rank's block of a row-partitioned CsrMatrix plus the NVLink peer connections.

    The constructor takes the rank's OWN rows (local row offsets from 0, global column indices,
    values) and the partition bounds; `from_global` slices them out of a global CSR, `grid3d`
    generates the slab of InitGrid3d directly so that no rank ever holds the global matrix."""

    def __init__(self, bounds, lro, ci_global, va, rank: int, world: int, all_gather_object):
        self.bounds = np.asarray(bounds, dtype=np.int64)
        self.rank, self.world = rank, world
        self.num_rows_global = int(self.bounds[-1])
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.plan = Plan(bounds, rank, world, self.num_rows_global, lro, ci_global, all_gather_object)
        self.n_local, self.n_halo = self.plan.n_local, self.plan.n_halo
        L = capi.lib()
        va = np.ascontiguousarray(va, dtype=np.float64)
        h = _P()
        capi._check(L.smle_dist_create_from_plan(C.byref(h), self.plan._h, va.ctypes.data_as(_P)))
        self._h = h
        mine = (C.c_ubyte * 64)()
        capi._check(L.smle_dist_ipc_handle(self._h, mine))
        handles = all_gather_object(bytes(mine))
        blob = b"".join(handles)
        capi._check(L.smle_dist_connect(self._h, blob))

    @classmethod
    def from_global(cls, ro, ci, va, rank: int, world: int, all_gather_object, bounds=None):
        ro = np.ascontiguousarray(ro, dtype=np.int32)
        if bounds is None:
            bounds = capi.dist_bounds(ro, world)   # merge-path search on the GPU, bit-exact with the reference
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        lo, hi = int(ro[r0]), int(ro[r1])
        lro = (ro[r0:r1 + 1].astype(np.int64) - lo).astype(np.int32)
        return cls(bounds, lro, ci[lo:hi], va[lo:hi], rank, world, all_gather_object)

    @classmethod
    def grid3d(cls, width: int, rank: int, world: int, all_gather_object, diag=6.0, offd=-1.0):
        """3-D Poisson slab of this rank: global row offsets (4 B per row) -> bounds -> rows [r0, r1) only."""
        ro = capi.gen_grid3d_row_offsets(width, True)
        bounds = capi.dist_bounds(ro, world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        nnz_rows = int(ro[r1]) - int(ro[r0])
        nnz_global = int(ro[-1])
        del ro
        lro, ci, va = capi.gen_grid3d_rows(width, r0, r1, nnz_rows, True, diag, offd)
        self = cls(bounds, lro, ci, va, rank, world, all_gather_object)
        self.num_nonzeros_global = nnz_global
        return self

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().smle_dist_destroy(self._h)
            self._h = None
            self.plan.close()

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
        """-> (iterations, x_local, final_rel_res); CGSolveSingle semantics on the global system.
        b_local / out: this rank's rows, torch CUDA tensors (device) or numpy / pinned CPU tensors (host)."""
        pb, dev, _ = capi._arg(b_local, np.float64)
        x = out if out is not None else capi._empty_like_arg(b_local, (self.n_local,), np.float64)
        px, dev_x, _ = capi._arg(x, np.float64, writable=True)
        if dev != dev_x:
            raise capi.SmleError("b and x must both be host or both be device memory")
        it, rel = _I(0), _D(0)
        capi._check(capi.lib().smle_dist_cg_f64(self._h, pb, px, _I(max_iters), _D(tolerance), _I(dev), C.byref(it), C.byref(rel)))
        return it.value, x, rel.value

    def cg_profile(self, b_local, iters: int):
        """mean ms of the three kernels of a row-partitioned iteration (CUDA events, no graph); collective"""
        pb, dev, _ = capi._arg(b_local, np.float64)
        if not dev:
            raise capi.SmleError("cg_profile needs a device tensor")
        out = (C.c_float * 3)()
        capi._check(capi.lib().smle_dist_cg_profile_f64(self._h, pb, _I(iters), out))
        return [float(v) for v in out]

    def allreduce_bench(self, iters: int = 6400) -> float:
        """microseconds per mailbox all-reduce of one double (measurement aid; collective)"""
        us = _D(0)
        capi._check(capi.lib().smle_dist_allreduce_bench_f64(self._h, _I(iters), C.byref(us)))
        return us.value


# ---------------------------------------------------------------------------------------------
# column-sharded multi-RHS CG (first row): the k recurrences of CGSolveMultiple
# never interact, so rank g solves columns [g*k/G, (g+1)*k/G) with A replicated; what the reference
# returns for the whole block follows from the shards' results alone
# ---------------------------------------------------------------------------------------------
def shard_columns(k: int, rank: int, world: int):
    """columns [lo, hi) of rank `rank` (equal shares, the first k % world ranks take one more)"""
    base, extra = divmod(k, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_sharded_results(results):
    """results: per rank (iterations, max_error_history) of CGSolveMultiple on its columns.
    -> (iterations, history) of the reference's lock-step solve of the whole block
    (no_pretreatment.hpp:133-161): it stops when EVERY column has latched, i.e. after the slowest shard's
    count; a shard that has stopped is frozen (alpha = beta = 0), so its columns keep contributing their
    final relative residual to the per-iteration maximum."""
    iters = max(r[0] for r in results)
    hist = np.zeros(iters, dtype=np.float64)
    for it, h in results:
        h = np.asarray(h, dtype=np.float64)
        if len(h) == 0:
            continue
        ext = np.concatenate([h[:iters], np.full(max(iters - len(h), 0), h[-1])])
        hist = np.maximum(hist, ext)
    return iters, hist


def cg_solve_multiple_column_sharded(a, B_local, max_iters: int, tolerance: float, all_gather_object, kernel_type: int = capi.MERGE):
    """One rank's part of the column-sharded CGSolveMultiple: `a` is this rank's (replicated) CsrMatrix handle,
    B_local its n x k_local columns.  -> (iterations of the whole block, X_local, history of the whole block).
    The only communication is one all-gather of (iterations, history) at the end."""
    it, X, hist, rel = a.cg_solve_multiple(B_local, max_iters, tolerance, kernel_type)
    iters, hist_all = merge_sharded_results(all_gather_object((it, hist)))
    return iters, X, hist_all
