// smle_spai.cpp -- host-side construction of the SPAI preconditioner.
//
// Input of the SPAI-preconditioned solve (smle_pcg_spai_multi_f64), not part of its hot loop: the
// reference builds M on the CPU as well (SparseApproximateInversion,
// work_2025/cg/sparse_approximate_inversion.hpp:41-321, called once per matrix at
// cpu_multicg.cpp:268-279).  Same definition: static pattern S_M = S_A; column k of M minimises
// || A(I,J) m - e_k(I) ||_2 with J = the rows of A's column k and I = the rows touched by the
// columns in J; the result is symmetrised, M <- (M + M^T)/2, entry by entry over the pattern.
// The reference solves every small problem with LAPACKE_?gels; here it is a column-major
// Householder QR written for the tiny, tall systems that occur (7 x 25 for a 7-point stencil).
#include "../../include/smle_b200.h"

#include <math.h>
#include <omp.h>

#include <vector>

namespace {

// min || Q x - y ||_2 for a column-major rows x cols matrix (rows >= cols); x overwrites y[0:cols].
// false: rank deficient (the reference then stores zeros for the column, :241-247).
bool least_squares(int rows, int cols, double *Q, double *y)
{
    if (rows < cols) return false;
    for (int j = 0; j < cols; ++j) {
        double *cj = Q + (size_t)j * rows;
        double tail = 0.0;
        for (int i = j + 1; i < rows; ++i) tail += cj[i] * cj[i];
        const double norm = sqrt(cj[j] * cj[j] + tail);
        if (norm == 0.0) return false;
        const double alpha = cj[j] > 0.0 ? -norm : norm;   // R(j,j)
        const double v0 = cj[j] - alpha;
        const double vv = v0 * v0 + tail;
        if (vv > 0.0) {
            auto reflect = [&](double *col) {
                double dot = v0 * col[j];
                for (int i = j + 1; i < rows; ++i) dot += cj[i] * col[i];
                const double f = 2.0 * dot / vv;
                col[j] -= f * v0;
                for (int i = j + 1; i < rows; ++i) col[i] -= f * cj[i];
            };
            for (int c = j + 1; c < cols; ++c) reflect(Q + (size_t)c * rows);
            reflect(y);
        }
        cj[j] = alpha;
    }
    for (int j = cols - 1; j >= 0; --j) {
        double s = y[j];
        for (int c = j + 1; c < cols; ++c) s -= Q[(size_t)c * rows + j] * y[c];
        y[j] = s / Q[(size_t)j * rows + j];
    }
    return true;
}

} // namespace

extern "C" int smle_spai_build_f64(int m, int nnz, const int *ro, const int *ci, const double *va, double *m_values)
{
    if (m < 0 || nnz < 0 || !ro || (nnz > 0 && (!ci || !va || !m_values))) return SMLE_ERR_ARG;
    // column view of A: for every column its (row, value, CSR position) triples in row order
    std::vector<int> cptr((size_t)m + 1, 0), crow((size_t)nnz), cpos((size_t)nnz);
    for (int z = 0; z < nnz; ++z) {
        if (ci[z] < 0 || ci[z] >= m) return SMLE_ERR_ARG;
        ++cptr[(size_t)ci[z] + 1];
    }
    for (int c = 0; c < m; ++c) cptr[(size_t)c + 1] += cptr[(size_t)c];
    {
        std::vector<int> fill(cptr.begin(), cptr.end() - 1);
        for (int r = 0; r < m; ++r)
            for (int z = ro[r]; z < ro[r + 1]; ++z) {
                const int d = fill[(size_t)ci[z]]++;
                crow[(size_t)d] = r;
                cpos[(size_t)d] = z;
            }
    }
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nnz; ++z) m_values[z] = 0.0;

#pragma omp parallel
    {
        std::vector<int> local_of((size_t)m, -1), I;
        std::vector<double> Q, y;
#pragma omp for schedule(dynamic, 256)
        for (int k = 0; k < m; ++k) {
            const int jb = cptr[(size_t)k], nv = cptr[(size_t)k + 1] - jb;
            if (nv == 0) continue;
            // I: rows of A that the columns in J reach, in first-seen order (the order only permutes the
            // equations of the least-squares problem)
            I.clear();
            for (int q = jb; q < jb + nv; ++q) {
                const int col = crow[(size_t)q];
                for (int t = cptr[(size_t)col]; t < cptr[(size_t)col + 1]; ++t) {
                    const int r = crow[(size_t)t];
                    if (local_of[(size_t)r] < 0) { local_of[(size_t)r] = (int)I.size(); I.push_back(r); }
                }
            }
            const int ne = (int)I.size();
            Q.assign((size_t)ne * nv, 0.0);
            y.assign((size_t)ne, 0.0);
            if (local_of[(size_t)k] >= 0) y[(size_t)local_of[(size_t)k]] = 1.0;
            for (int jl = 0; jl < nv; ++jl) {
                const int col = crow[(size_t)(jb + jl)];
                for (int t = cptr[(size_t)col]; t < cptr[(size_t)col + 1]; ++t)
                    Q[(size_t)jl * ne + local_of[(size_t)crow[(size_t)t]]] = va[cpos[(size_t)t]];
            }
            if (least_squares(ne, nv, Q.data(), y.data()))
                for (int jl = 0; jl < nv; ++jl) m_values[cpos[(size_t)(jb + jl)]] = y[(size_t)jl];   // M(J, k)
            for (int r : I) local_of[(size_t)r] = -1;
        }
    }
    // M <- (M + M^T) / 2 over the upper triangle of the pattern (:268-318)
#pragma omp parallel for schedule(static)
    for (int r = 0; r < m; ++r)
        for (int z = ro[r]; z < ro[r + 1]; ++z) {
            const int c = ci[z];
            if (c <= r) continue;
            for (int t = ro[c]; t < ro[c + 1]; ++t)
                if (ci[t] == r) {
                    const double avg = (m_values[z] + m_values[t]) * 0.5;
                    m_values[z] = avg;
                    m_values[t] = avg;
                    break;
                }
        }
    return SMLE_OK;
}
