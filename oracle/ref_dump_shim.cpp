// Test infrastructure only (oracle/). A dump shim compiled TOGETHER with the reference's
// own preproc.cpp so that tests can read the reference's internal per-graph tables
// (fields of `struct Preproc`, reference src/samplers/ugs_sampler/include/sampler.hpp:81-93)
// through the process-global `registry` (sampler.hpp:108-109). Plain C ABI, called via ctypes
// on the already-imported reference module. It only reads; nothing of the reference changes.
#include "sampler.hpp"
#include <cstring>

extern "C" {

// sizes: out[0]=n, out[1]=nnz, out[2]=alias.n ; returns 0 on success, -1 on unknown handle
int refdump_sizes(int64_t handle, int64_t* out) {
    std::lock_guard<std::mutex> lock(registry_mutex);
    auto it = registry.find(handle);
    if (it == registry.end()) return -1;
    auto& P = it->second;
    out[0] = P->n;
    out[1] = (int64_t)P->indices.size();
    out[2] = (int64_t)P->alias.n;
    return 0;
}

// Copies every table into caller-provided buffers (sizes from refdump_sizes).
int refdump_copy(int64_t handle,
                 int64_t* indptr,      // n+1
                 int32_t* indices,     // nnz
                 int32_t* edge_col,    // nnz
                 int32_t* order,       // n
                 int32_t* index_of,    // n
                 int32_t* suffix_deg,  // n
                 double*  bucket_b,    // n
                 double*  Z,           // 1
                 double*  alias_prob,  // alias.n
                 int32_t* alias_alias) // alias.n
{
    std::lock_guard<std::mutex> lock(registry_mutex);
    auto it = registry.find(handle);
    if (it == registry.end()) return -1;
    auto& P = it->second;
    const size_t n = (size_t)P->n, nnz = P->indices.size();
    std::memcpy(indptr, P->indptr.data(), (n + 1) * sizeof(int64_t));
    std::memcpy(indices, P->indices.data(), nnz * sizeof(int32_t));
    std::memcpy(edge_col, P->edge_col_of_csr_pos.data(), nnz * sizeof(int32_t));
    for (size_t i = 0; i < n; ++i) {
        order[i] = P->order[i];
        index_of[i] = P->index_of[i];
        suffix_deg[i] = P->suffix_deg[i];
        bucket_b[i] = P->bucket_b[i];
    }
    *Z = P->Z;
    for (int i = 0; i < P->alias.n; ++i) {
        alias_prob[i] = P->alias.prob[i];
        alias_alias[i] = P->alias.alias[i];
    }
    return 0;
}

}  // extern "C"
