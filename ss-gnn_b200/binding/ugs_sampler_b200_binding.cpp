// The pybind11/libtorch binding of the C ABI shown in INTEGRATION.md section 2, kept compilable:
// tests/test_gpu_binding.py builds it against libugs_b200.so and compares it with the ctypes drop-in.
// ugs_sampler_b200_binding.cpp  -- link with -lugs_b200
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include "ugs_b200.h"

static void chk(int rc) { TORCH_CHECK(rc == 0, ugs_last_error()); }

static py::tuple sample_batch_impl(const torch::Tensor& edge_index, const torch::Tensor& ptr,
                                   int m, int k, const std::string& mode, int seed, int fast) {
    TORCH_CHECK(edge_index.dtype() == torch::kInt64, "edge_index must be int64");
    TORCH_CHECK(ptr.dtype() == torch::kInt64, "ptr must be int64");
    TORCH_CHECK(mode == "sample" || mode == "graph" || mode == "global",
                "mode must be one of: 'sample', 'graph', 'global'");
    const int md = mode == "sample" ? UGS_MODE_SAMPLE : mode == "graph" ? UGS_MODE_GRAPH : UGS_MODE_GLOBAL;
    auto ptr_h = ptr.cpu();
    const int64_t G = ptr.size(0) - 1, E = edge_index.size(1);
    const int64_t base = G > 0 ? ptr_h[0].item<int64_t>() : 0, NT = G > 0 ? ptr_h[G].item<int64_t>() - base : 0;
    auto dev = torch::Device(torch::kCUDA);
    auto ei = edge_index.to(dev, /*non_blocking=*/true).contiguous();
    auto p = ptr.to(dev, true).contiguous();
    auto st = (void*)c10::cuda::getCurrentCUDAStream().stream();
    auto i64 = torch::TensorOptions().dtype(torch::kInt64).device(dev);
    auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(dev);

    ugs_layout L;  chk(ugs_workspace_layout(E, G, NT, &L));
    auto ws = torch::empty({(int64_t)L.total_bytes}, u8);
    int64_t info[16];
    chk(ugs_segment_batch(ei.data_ptr<int64_t>(), E, p.data_ptr<int64_t>(), G, k, 1, fast, nullptr, ws.data_ptr(), &L, st));
    chk(ugs_batch_info(ws.data_ptr(), &L, info, st));
    torch::Tensor perm;                      // only for batches whose columns are not graph-contiguous
    if (fast && (info[UGS_CTR_UNSORTED] & 4)) {  // not graph-contiguous: per-column keys, then a stable sort by graph
        chk(ugs_segment_batch(ei.data_ptr<int64_t>(), E, p.data_ptr<int64_t>(), G, k, 1, 0, nullptr, ws.data_ptr(), &L, st));
        chk(ugs_batch_info(ws.data_ptr(), &L, info, st));
        fast = 0;
    }
    if (info[UGS_CTR_UNSORTED] & 1) {
        const size_t sbytes = ugs_sort_scratch_bytes(E);
        auto sscratch = torch::empty({(int64_t)sbytes}, u8);
        perm = torch::empty({E}, i64);
        chk(ugs_sort_columns(E, G, ws.data_ptr(), &L, sscratch.data_ptr(), sbytes, perm.data_ptr<int64_t>(), st));
        chk(ugs_segment_batch(ei.data_ptr<int64_t>(), E, p.data_ptr<int64_t>(), G, k, 1, 0, perm.data_ptr<int64_t>(),
                              ws.data_ptr(), &L, st));
        chk(ugs_batch_info(ws.data_ptr(), &L, info, st));
    }
    const int64_t* pp = perm.defined() ? perm.data_ptr<int64_t>() : nullptr;
    chk(ugs_preproc_batch(ei.data_ptr<int64_t>(), E, p.data_ptr<int64_t>(), G, base, k, pp,
                          info[UGS_CTR_N_GENERAL], info[UGS_CTR_N_LARGE], ws.data_ptr(), &L, st));

    auto nodes = torch::empty({G * m, k}, i64), edge_ptr = torch::empty({G * m + 1}, i64),
         sample_ptr = torch::empty({G + 1}, i64);
    const size_t sb = ugs_sample_scratch_bytes(G, m);
    auto scratch = torch::empty({(int64_t)sb}, u8);
    chk(ugs_sample_batch(p.data_ptr<int64_t>(), G, base, m, k, (uint64_t)seed, /*graph_index_base=*/0,
                         /*sample_index_base=*/0, /*seed_mode=*/0, UGS_VARIANT_REF, info[UGS_CTR_N_GENERAL], info[UGS_CTR_MAX_N_GENERAL], ws.data_ptr(), &L,
                         scratch.data_ptr(), sb, nodes.data_ptr<int64_t>(), edge_ptr.data_ptr<int64_t>(),
                         sample_ptr.data_ptr<int64_t>(), st));
    chk(ugs_batch_info(ws.data_ptr(), &L, info, st));               // syncs; info[4] = total emitted entries
    if (fast && (info[UGS_CTR_UNSORTED] & 8))   // the preprocessing kernels found a column outside its
        return sample_batch_impl(edge_index, ptr, m, k, mode, seed, 0);  // graph's range: fast segmentation was fooled, redo
    const int64_t total = info[UGS_CTR_TOTAL_EDGES];
    auto edge_index_s = torch::empty({2, total}, i64), edge_src = torch::empty({total}, i64);
    chk(ugs_emit_edges(p.data_ptr<int64_t>(), G, base, m, k, md, perm.defined() ? 1 : 0, info[UGS_CTR_N_GENERAL], ws.data_ptr(), &L,
                       nodes.data_ptr<int64_t>(), edge_ptr.data_ptr<int64_t>(), total,
                       edge_index_s.data_ptr<int64_t>(), edge_src.data_ptr<int64_t>(), st));
    auto pin = [](const torch::Tensor& t) {                         // reference contract: CPU pinned outputs
        auto h = torch::empty(t.sizes(), t.options().device(torch::kCPU).pinned_memory(true));
        h.copy_(t, /*non_blocking=*/true);  return h; };
    auto out = py::make_tuple(pin(nodes), pin(edge_index_s), pin(edge_ptr), pin(sample_ptr), pin(edge_src));
    c10::cuda::getCurrentCUDAStream().synchronize();
    return out;
}

py::tuple sample_batch(const torch::Tensor& edge_index, const torch::Tensor& ptr,
                       int m, int k, const std::string& mode, int seed) {
    return sample_batch_impl(edge_index, ptr, m, k, mode, seed, /*fast=*/1);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("sample_batch", &sample_batch, py::arg("edge_index"), py::arg("ptr"), py::arg("m_per_graph"), py::arg("k"),
          py::arg("mode") = "sample", py::arg("seed") = 42,
          "Sample m_per_graph k-subgraphs per graph from a batched PyG edge_index + ptr (B200 C-ABI binding).");
}
