// smle_plan.h -- the partition plan object shared by smle_plan.cpp (which builds it, host only) and
// smle_capi.cu (which turns it into a device-resident row-partitioned system).  Internal.
#pragma once
#include <vector>

struct smle_plan_s {
    int rank = 0, world = 1, n_global = 0;
    int n_local = 0, n_halo = 0, halo_base = 0, nnz_local = 0;
    std::vector<int> bounds;      // world + 1 first rows
    std::vector<int> lro;         // local row offsets (n_local + 1)
    std::vector<int> lci;         // remapped column indices
    std::vector<int> halo_cols;   // global columns of the halo entries, ascending
    std::vector<int> need_off;    // world + 1: halo_cols[need_off[q] .. need_off[q+1]) are owned by rank q
    bool finished = false;
    std::vector<int> send_off, send_idx, send_dst, needs_from;
};

