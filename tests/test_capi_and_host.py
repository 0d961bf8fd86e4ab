"""CPU: the C-ABI library loads and exports every symbol include/ugs_b200.h declares; host-only entry points
work without a GPU; the drop-in modules keep the reference's signatures and fail loudly without CUDA."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ugs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ugs_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ugs_b200 import _lib
    so = _lib.build()
    L = ctypes.CDLL(so)
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ugs_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    assert L.ugs_abi_version() == 3


def test_workspace_layout_is_host_only_and_consistent():
    from ugs_b200 import _lib
    L = _lib.lib()
    lay = _lib.Layout()
    _lib.check(L.ugs_workspace_layout(1000, 10, 300, ctypes.byref(lay)))
    offs = [getattr(lay, f) for f, _ in _lib.Layout._fields_[1:]]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    assert lay.total_bytes > offs[-1]
    assert lay.csr_col - lay.csr_idx >= 4 * 2 * 1000
    with pytest.raises(RuntimeError, match="too large"):
        _lib.check(L.ugs_workspace_layout(1 << 31, 10, 300, ctypes.byref(lay)))
    assert L.ugs_sample_scratch_bytes(10, 64) >= 16


def test_dropin_signatures_match_the_reference_stubs():
    """Argument names/defaults are part of the ABI (keyword callers: gps/utils/wl_vocab.py:144-151)."""
    import ugs_sampler as us
    sig = inspect.signature(us.sample_batch)
    assert list(sig.parameters) == ["edge_index", "ptr", "m_per_graph", "k", "mode", "seed"]
    assert sig.parameters["mode"].default == "sample" and sig.parameters["seed"].default == 42
    sig = inspect.signature(us.sample)
    assert list(sig.parameters) == ["handle", "m_per_graph", "k", "edge_mode", "base_offset", "seed"]
    assert sig.parameters["edge_mode"].default == "local" and sig.parameters["base_offset"].default == 0
    assert list(inspect.signature(us.create_preproc).parameters) == ["edge_index", "num_nodes", "k"]
    for f in ("destroy_preproc", "has_graphlets", "get_preproc_info"):
        assert list(inspect.signature(getattr(us, f)).parameters) == ["handle"]
    assert us.get_preproc_info(12345) == {} and us.has_graphlets(12345) is False
    us.destroy_preproc(12345)  # unknown handles are ignored
    with pytest.raises(RuntimeError, match="Invalid preproc handle"):
        us.sample(12345, 1, 3)


def test_all_five_sampler_modules_keep_the_reference_signatures():
    """src/samplers/*/__init__.pyi and the py::arg lists of the five pybind11 modules (rwr_sampler.cpp:297-308,
    uniform_sampler.cpp:287-295, epsilon_uniform_sampler.cpp:367-377, apx_ugs_sampler.cpp:529-538)."""
    import apx_ugs_sampler
    import epsilon_uniform_sampler
    import rwr_sampler
    import uniform_sampler
    base = ["edge_index", "ptr", "m_per_graph", "k", "mode", "seed"]
    for mod, extra in ((rwr_sampler, [("p_restart", 0.2)]), (uniform_sampler, []),
                       (epsilon_uniform_sampler, [("epsilon", 0.1)]), (apx_ugs_sampler, [("epsilon", 0.1)])):
        sig = inspect.signature(mod.sample_batch)
        public = [n for n in sig.parameters if not n.startswith("_")]  # apx has a private test hook (_max_trials)
        assert public == base + [n for n, _ in extra], mod.__name__
        assert sig.parameters["mode"].default == "sample" and sig.parameters["seed"].default == 42
        for n, d in extra:
            assert sig.parameters[n].default == d


def test_host_side_modules_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ugs_b200 import encode, presample, wl
    x = torch.zeros(4, 2)
    idx = torch.zeros(3, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU path"):
        encode.gather_rows(x, idx)
    with pytest.raises(RuntimeError, match="no CPU path"):
        wl.wl_digests(torch.zeros((1, 3), dtype=torch.long), torch.zeros((2, 0), dtype=torch.long),
                      torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU path"):
        presample.PresampleStore.build(torch.zeros((2, 0), dtype=torch.long), torch.zeros(1, dtype=torch.long), 2, 3)
    import rwr_sampler
    import uniform_sampler
    for mod in (rwr_sampler, uniform_sampler):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            mod.sample_batch(torch.tensor([[0, 1], [1, 2]]), torch.tensor([0, 3]), 2, 2)


def test_dropin_argument_errors_and_no_cpu_fallback():
    import torch
    import ugs_sampler as us
    ei = torch.tensor([[0, 1], [1, 2]])
    ptr = torch.tensor([0, 3])
    with pytest.raises(RuntimeError, match="int64"):
        us.sample_batch(ei.to(torch.int32), ptr, 2, 2)
    with pytest.raises(RuntimeError, match="ptr must be int64"):
        us.sample_batch(ei, ptr.to(torch.int32), 2, 2)
    with pytest.raises(RuntimeError, match="mode must be one of"):
        us.sample_batch(ei, ptr, 2, 2, mode="bogus")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            us.sample_batch(ei, ptr, 2, 2)


def test_product_never_imports_the_oracle():
    """The product path must not import, include, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "ss-gnn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            txt = open(os.path.join(dp, f)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
            assert not re.search(r"#include[^\n]*oracle", txt), f
            assert "libugs_oracle" not in txt and "ref_loader" not in txt, f


def test_synthetic_workloads_have_the_named_shapes():
    from ugs_b200 import synth
    ei, ptr = synth.make_batch("zinc", 500, seed=1)
    n = np.diff(ptr)
    assert n.min() >= 9 and n.max() <= 38
    assert (ei[0] != ei[1]).all()
    # bidirectional and graph-contiguous
    key = set(map(tuple, ei.T.tolist()))
    assert all((b, a) in key for a, b in key)
    g0 = np.searchsorted(ptr, ei[0], side="right") - 1
    g1 = np.searchsorted(ptr, ei[1], side="right") - 1
    assert (g0 == g1).all() and (np.diff(g0) >= 0).all()
    assert 1.9 < ei.shape[1] / ptr[-1] < 2.4  # average degree ~2.15
    ei, ptr = synth.make_batch("proteins", 400, seed=2)
    assert np.diff(ptr).max() <= 620 and np.diff(ptr).min() >= 4
    ei, ptr = synth.barabasi_albert(500, 3, seed=3)
    deg = np.bincount(ei[0], minlength=500)
    assert deg.max() > 20 and ptr.tolist() == [0, 500]


def test_ugs_debug_lines_have_the_reference_formats():
    """UGS_DEBUG=1 diagnostics (reference preproc.cpp:253-254, sampler.cpp:184-186,
    ugs_sampler_batch_extension.cpp:249-254): same prefixes, field names, order and number formats."""
    import ugs_sampler as us
    assert us._fmt_preproc(23, 8, 1.5e7, 17) == "[UGS PREPROC] n=23 k=8 Z=1.50e+07 viable=17/23"
    assert (us._fmt_sample(23, 8, 64, 1.5e7, 0, 0)
            == "[UGS SAMPLE] n=23 k=8 m=64 Z=1.50e+07 | relax=0 weighted=64 uniform=0 incomplete=0")
    assert (us._fmt_sample(3, 3, 5, 0.0, 1, 5)
            == "[UGS SAMPLE] n=3 k=3 m=5 Z=0.00e+00 | relax=1 weighted=0 uniform=5 incomplete=5")
    line = us._fmt_cache(10)
    assert re.match(r"^\[UGS CACHE\] hits=0 misses=10 hit_rate=0\.0% cache_size=0 \(cache: none", line)
    os.environ.pop("UGS_DEBUG", None)
    assert not us._debug_on()
    os.environ["UGS_DEBUG"] = "true"  # the reference compares with the string "1" only (preproc.cpp:247)
    assert not us._debug_on()
    os.environ["UGS_DEBUG"] = "1"
    assert us._debug_on()
    del os.environ["UGS_DEBUG"]


def test_host_widening_variants_equal_numpy():
    """csrc/ugs_hostdeliver.cpp: every widening variant this CPU has (scalar / AVX2 / AVX-512) against numpy, at odd
    lengths and destination alignments (row 1 of edge_index starts wherever Ebar puts it)."""
    from ugs_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    ran = set()
    for isa in (b"scalar", b"avx2", b"avx512"):
        for n in (0, 1, 7, 8, 9, 63, 64, 1000, 4099):
            for shift in (0, 1, 3, 5):
                for kind in (0, 1, 2, 3, 4):
                    if kind == 4:
                        src = rng.integers(0, 256, size=n, dtype=np.int64).astype(np.uint8)
                        if n:
                            src[rng.integers(0, n, size=max(1, n // 5))] = 0xFF
                        want = np.where(src == 0xFF, -1, 10 ** 12 + src.astype(np.int64))
                    elif kind == 0:
                        src = rng.integers(-2 ** 31, 2 ** 31, size=n, dtype=np.int64).astype(np.int32)
                        want = src.astype(np.int64)
                    elif kind in (1, 2):
                        src = rng.integers(0, 256, size=n, dtype=np.int64).astype(np.uint8)
                        want = (src & 15).astype(np.int64) if kind == 1 else (src >> 4).astype(np.int64)
                    else:
                        src = rng.integers(0, 65536, size=n, dtype=np.int64).astype(np.uint16)
                        if n:
                            src[rng.integers(0, n, size=max(1, n // 5))] = 0xFFFF
                        want = np.where(src == 0xFFFF, -1, 10 ** 12 + src.astype(np.int64))
                    buf = np.full(n + 16, -7, np.int64)
                    dst = buf[shift:shift + n]
                    rc = L.ugs_host_widen(kind, src.ctypes.data if n else None, dst.ctypes.data, n,
                                          10 ** 12 if kind >= 3 else 0, isa)
                    if rc == -1:
                        continue
                    assert rc == 0
                    ran.add(isa)
                    assert np.array_equal(dst, want), (isa, n, shift, kind)
                    assert (buf[:shift] == -7).all() and (buf[shift + n:] == -7).all()
    assert b"scalar" in ran
    assert L.ugs_host_isa().decode() in ("scalar", "avx2", "avx512")


def test_host_copy_is_a_parallel_memcpy():
    """ugs_host_copy (pageable input -> pinned staging buffer, csrc/ugs_hostdeliver.cpp): every byte, at sizes around the
    per-thread slice boundaries, with one thread and with more threads than slices."""
    from ugs_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    for n in (0, 1, 4095, 4096, 4097, (1 << 20) - 1, (1 << 20) + 1, 5 * (1 << 20) + 123):
        src = rng.integers(0, 256, size=n, dtype=np.uint8)
        for threads in (1, 3, 16):
            buf = np.full(n + 64, 7, np.uint8)
            assert L.ugs_host_copy(buf[32:].ctypes.data, src.ctypes.data if n else buf.ctypes.data, n, threads) == 0
            assert np.array_equal(buf[32:32 + n], src) and (buf[:32] == 7).all() and (buf[32 + n:] == 7).all()


def test_host_store_bandwidth_probe_and_pool():
    """The host-side roofline probe runs on the persistent worker pool (no GPU needed) and is re-entrant from threads."""
    import ctypes
    import threading
    from ugs_b200 import _lib
    L = _lib.lib()
    buf = np.zeros(1 << 22, np.int64)
    res = []

    def run():
        res.append(float(L.ugs_host_store_bandwidth(ctypes.c_void_p(buf.ctypes.data), buf.nbytes, 4, 2)))

    ts = [threading.Thread(target=run) for _ in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(res) == 3 and all(r > 0.1 for r in res)
    assert (buf == 1).any() or (buf == 0).any()  # written with the probe's pattern (low nibbles of its source block)
