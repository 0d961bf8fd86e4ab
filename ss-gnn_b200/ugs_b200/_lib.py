"""ctypes binding of libugs_b200.so (the C ABI declared in include/ugs_b200.h) and its nvcc build."""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SO = os.environ.get("UGS_B200_LIB") or os.path.join(HERE, "libugs_b200.so")  # override: A/B kernel variants
SOURCES = ["ugs_preproc.cu", "ugs_preproc_large.cu", "ugs_preproc_medium.cu", "ugs_sortcols.cu", "ugs_sample.cu", "ugs_gather.cu", "ugs_rwr.cu", "ugs_enum.cu", "ugs_presample.cu", "ugs_wl.cu", "ugs_hostio.cu", "ugs_hostdeliver.cpp", "ugs_fused.cu", "ugs_variants.cu", "ugs_uniform_general.cu", "ugs_apx.cu", "ugs_capi.cu"]
HEADERS = [os.path.join(CSRC, "ugs_common.cuh"), os.path.join(CSRC, "ugs_hostio.h"), os.path.join(REPO, "include", "ugs_b200.h")]

EXPORTS = [
    "ugs_abi_version", "ugs_launch_count", "ugs_last_error", "ugs_workspace_layout", "ugs_segment_batch", "ugs_batch_info",
    "ugs_preproc_batch", "ugs_sample_scratch_bytes", "ugs_sample_batch", "ugs_emit_edges", "ugs_count_edges",
    "ugs_eps_sample_batch", "ugs_eps_emit_edges", "ugs_preproc_uniform",
    "ugs_sort_scratch_bytes", "ugs_sort_columns", "ugs_apx_workspace_bytes", "ugs_apx_prepare", "ugs_apx_trials", "ugs_apx_write", "ugs_apx_dump",
    "ugs_gather_rows", "ugs_encoder_offsets", "ugs_rwr_sample_batch", "ugs_rwr_emit_edges",
    "ugs_enum_scratch_bytes", "ugs_enum_record_words", "ugs_enum_count", "ugs_enum_fill", "ugs_enum_sample_batch",
    "ugs_presample_scratch_bytes", "ugs_presample_plan", "ugs_presample_load",
    "ugs_wl_hash", "ugs_wl_lookup", "ugs_wl_node_attr",
    "ugs_batch_info_begin", "ugs_batch_info_end", "ugs_preproc_batch_parts",
    "ugs_deliver_wire_bytes", "ugs_deliver_outputs", "ugs_host_store_bandwidth", "ugs_host_isa", "ugs_host_widen",
    "ugs_emit_edges_packed", "ugs_sample_batch_early", "ugs_emit_edges_wire", "ugs_deliver_outputs_wire", "ugs_host_copy",
    "ugs_fused_capacity", "ugs_sample_batch_fused", "ugs_fused_finish",
]


class LibraryMissing(RuntimeError):
    pass


class Layout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in (
        "total_bytes", "cstart", "ginfo", "glist", "counters", "colkey", "rowptr", "order", "index_of",
        "suffix_deg", "bucket_b", "cum", "adj", "suf", "csr_idx", "csr_col", "gaux", "llist", "lscratch")]


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", f"-I{os.path.join(REPO, 'include')}", f"-I{CSRC}"]


def build(force: bool = False, extra_flags=(), out: str = SO) -> str:
    """Compile libugs_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU): one object per source file
    under csrc/build/ (only stale ones are recompiled, in parallel), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    bdir = os.path.join(CSRC, "build" if out == SO else "build_" + os.path.basename(out))
    os.makedirs(bdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in HEADERS + [os.path.abspath(__file__)])
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(bdir, os.path.splitext(os.path.basename(s))[0] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            if s.endswith(".cpp"):  # host-only code (thread pool, AVX widening): the host compiler, no nvcc front end
                jobs.append(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", f"-I{os.path.join(REPO, 'include')}",
                             f"-I{CSRC}", "-I/usr/local/cuda/include", "-c", s, "-o", o])
            else:
                jobs.append(["nvcc"] + NVCC_FLAGS + list(extra_flags) + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(subprocess.check_call, jobs))
    if jobs or not os.path.exists(out) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
                              + objs + ["-o", out])
    return out


_LIB = None


def lib():
    """Load the shared library. There is no CPU fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO):
        raise LibraryMissing(
            f"{SO} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The B200 sampler has no CPU fallback.")
    L = ctypes.CDLL(SO)
    vp, i64, u64, sz, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    LP = ctypes.POINTER(Layout)
    L.ugs_abi_version.restype = ci
    L.ugs_last_error.restype = ctypes.c_char_p
    L.ugs_launch_count.restype = ctypes.c_uint64
    L.ugs_workspace_layout.argtypes = [i64, i64, i64, LP]
    L.ugs_segment_batch.argtypes = [vp, i64, vp, i64, ci, ci, ci, vp, vp, LP, vp]
    L.ugs_batch_info.argtypes = [vp, LP, vp, vp]
    L.ugs_preproc_batch.argtypes = [vp, i64, vp, i64, i64, ci, vp, i64, i64, vp, LP, vp]
    L.ugs_sample_scratch_bytes.restype = sz
    L.ugs_sample_scratch_bytes.argtypes = [i64, i64]
    L.ugs_sample_batch.argtypes = [vp, i64, i64, i64, ci, u64, u64, u64, ci, ci, i64, i64, vp, LP, vp, sz, vp, vp, vp, vp]
    L.ugs_emit_edges.argtypes = [vp, i64, i64, i64, ci, ci, ci, i64, vp, LP, vp, vp, i64, vp, vp, vp]
    L.ugs_count_edges.argtypes = [vp, i64, i64, i64, ci, i64, vp, LP, vp, sz, vp, vp, vp]
    cd = ctypes.c_double
    L.ugs_apx_workspace_bytes.restype = sz
    L.ugs_apx_workspace_bytes.argtypes = [i64, i64]
    L.ugs_apx_prepare.argtypes = [vp, i64, i64, i64, i64, ci, u64, cd, vp, sz, vp, vp]
    L.ugs_apx_trials.argtypes = [ci, i64, i64, ci, u64, cd, ci, ci, ci, vp, vp, vp]
    L.ugs_apx_write.argtypes = [ci, i64, i64, ci, u64, cd, vp, ci, vp, vp, vp]
    L.ugs_apx_dump.argtypes = [ci, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.ugs_sort_scratch_bytes.restype = sz
    L.ugs_sort_scratch_bytes.argtypes = [i64]
    L.ugs_sort_columns.argtypes = [i64, i64, vp, LP, vp, sz, vp, vp]
    L.ugs_preproc_uniform.argtypes = [vp, i64, i64, ci, i64, i64, vp, LP, vp]
    L.ugs_eps_sample_batch.argtypes = [vp, i64, i64, i64, ci, u64, u64, ctypes.c_double, i64, vp, LP, vp, sz, vp, vp, vp, vp]
    L.ugs_eps_emit_edges.argtypes = [vp, i64, vp, i64, i64, i64, ci, ci, vp, i64, vp, LP, vp, vp, i64, vp, vp, vp]
    L.ugs_gather_rows.argtypes = [vp, i64, i64, vp, i64, vp, vp, vp]
    L.ugs_encoder_offsets.argtypes = [vp, vp, i64, ci, i64, vp, vp, vp]
    L.ugs_rwr_sample_batch.argtypes = [vp, i64, i64, i64, ci, u64, u64, cd, vp, LP, vp, sz, vp, vp, vp, vp]
    L.ugs_rwr_emit_edges.argtypes = L.ugs_emit_edges.argtypes
    L.ugs_enum_scratch_bytes.restype = sz
    L.ugs_enum_scratch_bytes.argtypes = [i64, i64, i64, i64]
    L.ugs_enum_record_words.argtypes = [i64, i64]
    L.ugs_enum_count.argtypes = [vp, i64, i64, i64, ci, i64, i64, vp, LP, vp, sz, vp]
    L.ugs_enum_fill.argtypes = [vp, i64, i64, i64, ci, i64, i64, vp, LP, vp, vp, i64, vp]
    L.ugs_enum_sample_batch.argtypes = [vp, i64, i64, i64, i64, ci, u64, u64, i64, i64, vp, LP, vp, vp, vp, sz, vp, vp,
                                        vp, vp]
    L.ugs_presample_scratch_bytes.restype = sz
    L.ugs_presample_scratch_bytes.argtypes = [i64]
    L.ugs_presample_plan.argtypes = [vp, i64, i64, vp, i64, vp, vp, i64, vp, sz, ctypes.POINTER(i64), vp]
    L.ugs_presample_load.argtypes = [vp, vp, i64, vp, vp, vp, vp, i64, i64, ci, vp, i64, vp, vp, i64, ci,
                                     vp, vp, vp, vp, vp, vp]
    L.ugs_wl_hash.argtypes = [vp, i64, ci, vp, i64, vp, ci, vp, i64, vp, vp, vp]
    L.ugs_wl_node_attr.argtypes = [vp, i64, i64, vp, vp]
    L.ugs_wl_lookup.argtypes = [vp, vp, i64, vp, vp, i64, i64, vp, vp]
    L.ugs_batch_info_begin.argtypes = [vp, LP, vp, vp]
    L.ugs_batch_info_end.argtypes = []
    L.ugs_preproc_batch_parts.argtypes = [vp, i64, vp, i64, i64, ci, vp, i64, i64, ci, vp, LP, vp]
    L.ugs_emit_edges_packed.argtypes = [vp, i64, i64, i64, ci, ci, ci, i64, vp, LP, vp, vp, i64, vp, vp, vp, sz, vp]
    L.ugs_sample_batch_early.argtypes = [vp, i64, i64, i64, ci, u64, u64, u64, ci, ci, i64, i64, vp, LP, vp, sz, vp, vp, vp, vp, vp]
    L.ugs_fused_capacity.restype = i64
    L.ugs_fused_capacity.argtypes = [i64, i64]
    L.ugs_sample_batch_fused.argtypes = [vp, i64, vp, i64, i64, i64, i64, ci, ci, u64, u64, u64, ci, vp, LP, vp, sz,
                                         vp, vp, vp, vp, vp, vp, i64, vp, ci, vp]
    L.ugs_fused_finish.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    L.ugs_deliver_wire_bytes.restype = sz
    L.ugs_deliver_wire_bytes.argtypes = [i64, ci, i64, ci]
    L.ugs_deliver_outputs.argtypes = [vp, i64, ci, vp, vp, i64, ci, cd, vp, vp, sz, vp, vp, vp, ci, vp, vp]
    L.ugs_emit_edges_wire.argtypes = [vp, i64, i64, i64, ci, i64, vp, LP, vp, vp, i64, vp, sz, vp, sz, vp]
    L.ugs_deliver_outputs_wire.argtypes = [vp, i64, ci, i64, vp, vp, sz, vp, vp, vp, ci, vp, vp]
    L.ugs_host_copy.argtypes = [vp, vp, sz, ci]
    L.ugs_host_store_bandwidth.restype = cd
    L.ugs_host_store_bandwidth.argtypes = [vp, sz, ci, ci]
    L.ugs_host_isa.restype = ctypes.c_char_p
    L.ugs_host_widen.argtypes = [ci, vp, vp, sz, i64, ctypes.c_char_p]
    _LIB = L
    return L


def check(rc: int):
    if rc != 0:
        msg = lib().ugs_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"libugs_b200 error {rc}")
