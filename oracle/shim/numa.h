/* oracle/shim/numa.h -- stand-in for libnuma (not installed in this image) so the
 * UNMODIFIED reference headers compile.  numa_available() < 0 makes
 * CsrMatrix::IsNumaMalloc() false (sparse_matrix.h:656-663), selecting the
 * mkl_malloc branch (:699-704).  Test infrastructure only. */
#ifndef SMLE_SHIM_NUMA_H
#define SMLE_SHIM_NUMA_H
#include <stdlib.h>
static inline int numa_available(void) { return -1; }
static inline void numa_set_strict(int) {}
static inline int numa_num_task_nodes(void) { return 1; }
static inline void *numa_alloc_onnode(size_t bytes, int) { return malloc(bytes); }
static inline void numa_free(void *p, size_t) { free(p); }
#endif
