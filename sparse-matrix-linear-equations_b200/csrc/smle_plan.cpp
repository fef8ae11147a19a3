// smle_plan.cpp -- host-side planner of the row-partitioned path.
//
// Net-new relative to the reference (no distributed code there).  Given the rows a rank owns --
// cut at the reference's merge-path coordinates, x_g = MergePathSearch(min(g*ceil((m+nnz)/G),
// m+nnz)).x (work_2025/spmm/merge_based.hpp:22-44, share diagonals :72-82) -- the planner builds
//   * the local system: the rank's rows with columns remapped to [own | pad | halo], halo = the
//     sorted unique out-of-range columns (the halo index map), starting on a 128-byte line of its
//     own so that no cache line holds both local entries and entries a peer writes;
//   * the request every rank publishes (which global columns it needs, grouped by owner);
//   * from all ranks' requests, the push plan: which local rows go to which peer and where they
//     land in that peer's extended vector.
// Nothing here touches a GPU and no rank ever needs more than its own rows: the only exchange is
// one all-gather of the request blobs, moved by the caller (torch.distributed, or the shared-memory
// segment of the multi-process C++ driver).
#include "../../include/smle_b200.h"

#include <algorithm>
#include <new>

#include "smle_plan.h"

namespace {
constexpr int kHaloAlign = 16;    // doubles per 128-byte line
constexpr int kBlobHeader = 3;    // n_local, n_halo, halo_base, then need_off[world+1], then halo_cols
}

extern "C" {

int smle_dist_plan_create(smle_plan_t *out, int rank, int world, const int *bounds, int num_cols_global,
                          const int *local_row_offsets, const int *global_column_indices)
{
    if (!out || world < 1 || rank < 0 || rank >= world || !bounds || !local_row_offsets || num_cols_global < 0)
        return SMLE_ERR_ARG;
    for (int q = 0; q < world; ++q)
        if (bounds[q] > bounds[q + 1]) return SMLE_ERR_ARG;
    if (bounds[0] != 0 || bounds[world] != num_cols_global) return SMLE_ERR_ARG;   // square system: every column has an owner
    smle_plan_s *p = new (std::nothrow) smle_plan_s();
    if (!p) return SMLE_ERR_ALLOC;
    p->rank = rank; p->world = world; p->n_global = num_cols_global;
    p->bounds.assign(bounds, bounds + world + 1);
    const int r0 = bounds[rank], r1 = bounds[rank + 1];
    p->n_local = r1 - r0;
    p->lro.assign(local_row_offsets, local_row_offsets + p->n_local + 1);
    if (p->lro[0] != 0) { delete p; return SMLE_ERR_ARG; }
    p->nnz_local = p->lro[p->n_local];
    if (p->nnz_local > 0 && !global_column_indices) { delete p; return SMLE_ERR_ARG; }
    const int *ci = global_column_indices;
    const long long nnz = p->nnz_local;

    // halo index map: sorted unique columns outside [r0, r1)
    std::vector<int> &halo = p->halo_cols;
    for (long long z = 0; z < nnz; ++z) {
        const int c = ci[z];
        if (c < 0 || c >= num_cols_global) { delete p; return SMLE_ERR_ARG; }
        if (c < r0 || c >= r1) halo.push_back(c);
    }
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    p->n_halo = (int)halo.size();
    p->halo_base = (p->n_local + kHaloAlign - 1) / kHaloAlign * kHaloAlign;
    if ((long long)p->halo_base + p->n_halo > 2147483647LL) { delete p; return SMLE_ERR_RANGE; }

    p->need_off.assign((size_t)world + 1, 0);
    for (int q = 0; q <= world; ++q)
        p->need_off[q] = (int)(std::lower_bound(halo.begin(), halo.end(), p->bounds[q]) - halo.begin());
    p->need_off[world] = p->n_halo;

    p->lci.resize((size_t)nnz);
    const int hb = p->halo_base;
#pragma omp parallel for schedule(static)
    for (long long z = 0; z < nnz; ++z) {
        const int c = ci[z];
        p->lci[(size_t)z] = (c >= r0 && c < r1) ? c - r0
                                                 : hb + (int)(std::lower_bound(halo.begin(), halo.end(), c) - halo.begin());
    }
    *out = p;
    return SMLE_OK;
}

void smle_dist_plan_destroy(smle_plan_t p) { delete p; }

int smle_dist_plan_dims(smle_plan_t p, int *n_local, int *n_halo, int *halo_base, int *nnz_local)
{
    if (!p) return SMLE_ERR_ARG;
    if (n_local) *n_local = p->n_local;
    if (n_halo) *n_halo = p->n_halo;
    if (halo_base) *halo_base = p->halo_base;
    if (nnz_local) *nnz_local = p->nnz_local;
    return SMLE_OK;
}

int smle_dist_plan_local_columns(smle_plan_t p, int *out)
{
    if (!p || (!out && p->nnz_local)) return SMLE_ERR_ARG;
    std::copy(p->lci.begin(), p->lci.end(), out);
    return SMLE_OK;
}

int smle_dist_plan_halo_columns(smle_plan_t p, int *out)
{
    if (!p || (!out && p->n_halo)) return SMLE_ERR_ARG;
    std::copy(p->halo_cols.begin(), p->halo_cols.end(), out);
    return SMLE_OK;
}

long long smle_dist_plan_request_size(smle_plan_t p)
{
    if (!p) return SMLE_ERR_ARG;
    return (long long)kBlobHeader + p->world + 1 + p->n_halo;
}

int smle_dist_plan_request(smle_plan_t p, int *blob)
{
    if (!p || !blob) return SMLE_ERR_ARG;
    blob[0] = p->n_local; blob[1] = p->n_halo; blob[2] = p->halo_base;
    std::copy(p->need_off.begin(), p->need_off.end(), blob + kBlobHeader);
    std::copy(p->halo_cols.begin(), p->halo_cols.end(), blob + kBlobHeader + p->world + 1);
    return SMLE_OK;
}

int smle_dist_plan_finish(smle_plan_t p, const int *all_blobs, const long long *blob_off)
{
    if (!p || !all_blobs || !blob_off) return SMLE_ERR_ARG;
    const int world = p->world, rank = p->rank, r0 = p->bounds[rank];
    p->send_off.assign((size_t)world + 1, 0);
    p->send_dst.assign((size_t)world, 0);
    p->needs_from.assign((size_t)world, 0);
    p->send_idx.clear();
    for (int q = 0; q < world; ++q) {
        const int *b = all_blobs + blob_off[q];
        const long long len = blob_off[q + 1] - blob_off[q];
        if (len < kBlobHeader + world + 1) return SMLE_ERR_ARG;
        const int q_n_local = b[0], q_n_halo = b[1], q_halo_base = b[2];
        const int *q_need_off = b + kBlobHeader, *q_cols = b + kBlobHeader + world + 1;
        if (len != (long long)kBlobHeader + world + 1 + q_n_halo || q_n_local != p->bounds[q + 1] - p->bounds[q])
            return SMLE_ERR_ARG;
        if (q != rank) {
            const int lo = q_need_off[rank], hi = q_need_off[rank + 1];
            if (lo < 0 || hi < lo || hi > q_n_halo) return SMLE_ERR_ARG;
            for (int i = lo; i < hi; ++i) {
                const int li = q_cols[i] - r0;
                if (li < 0 || li >= p->n_local) return SMLE_ERR_ARG;   // q asks this rank for a row it does not own
                p->send_idx.push_back(li);
            }
            p->send_dst[q] = q_halo_base + lo;   // where the group lands in q's extended vector
            p->needs_from[q] = p->need_off[q + 1] > p->need_off[q] ? 1 : 0;
        }
        p->send_off[q + 1] = (int)p->send_idx.size();
    }
    p->finished = true;
    return SMLE_OK;
}

int smle_dist_plan_send(smle_plan_t p, int *send_off, int *send_idx, int *send_dst, int *needs_from)
{
    if (!p || !p->finished) return SMLE_ERR_ARG;
    if (send_off) std::copy(p->send_off.begin(), p->send_off.end(), send_off);
    if (send_idx) std::copy(p->send_idx.begin(), p->send_idx.end(), send_idx);
    if (send_dst) std::copy(p->send_dst.begin(), p->send_dst.end(), send_dst);
    if (needs_from) std::copy(p->needs_from.begin(), p->needs_from.end(), needs_from);
    return SMLE_OK;
}

} // extern "C"
