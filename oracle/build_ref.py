"""Test infrastructure only (oracle/): build the UNMODIFIED reference samplers into oracle/_ref/.

The reference C++ (pybind11 + libtorch, no CUDA) is compiled from the sources WHERE THEY LIE under
/root/reference/src/samplers/* with g++ directly (no reference build system, no source copies):

  oracle/_ref/pinned/ugs_sampler.so              reference ugs_sampler + oracle/ref_dump_shim.cpp
  oracle/_ref/nopin/ugs_sampler.so               same, force-including oracle/ref_nopin_prefix.h so
                                                 it can run on a host without an NVIDIA driver
  oracle/_ref/{pinned,nopin}/epsilon_uniform_sampler.so
  oracle/_ref/{pinned,nopin}/apx_ugs_sampler.so
  oracle/_ref/{pinned,nopin}/uniform_sampler.so
  oracle/_ref/{pinned,nopin}/rwr_sampler.so

`oracle/_ref/` is git-ignored but NOT gpurun-ignored: the built files travel to the GPU box, where
/root/reference does not exist. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may load these files.

Recipe follows SURVEY.md §8(c)/Appendix B (absolute -I, CXX=/usr/bin/g++ for libgomp).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SSGNN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _flags():
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths() + [pybind11.get_include(), sysconfig.get_paths()["include"]]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cflags = ["-O3", "-std=c++17", "-fPIC", "-shared", "-w",
              "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cflags += [f"-I{p}" for p in inc]
    ldflags = [f"-L{tlib}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", f"-Wl,-rpath,{tlib}"]
    return cflags, ldflags


def _compile(name, sources, include_dir, variant, extra_c=(), extra_ld=()):
    cflags, ldflags = _flags()
    out_dir = os.path.join(OUT, variant)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    newest = max(os.path.getmtime(s) for s in sources)
    if os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    cmd = [CXX] + cflags + list(extra_c) + [f"-I{include_dir}", f"-DTORCH_EXTENSION_NAME={name}"]
    if variant == "nopin":
        cmd += ["-include", os.path.join(HERE, "ref_nopin_prefix.h")]
    cmd += sources + ["-o", out] + ldflags + list(extra_ld)
    print("[oracle/_ref]", variant, name, file=sys.stderr)
    subprocess.check_call(cmd)
    return out


def build_all(variants=("pinned", "nopin")):
    """Build every reference sampler on the path. Returns dict {(variant, name): path}."""
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"reference tree not found at {REF}")
    S = os.path.join(REF, "src", "samplers")
    built = {}
    for v in variants:
        ugs = os.path.join(S, "ugs_sampler")
        built[(v, "ugs_sampler")] = _compile(
            "ugs_sampler",
            [os.path.join(ugs, "src", f) for f in
             ("extension.cpp", "preproc.cpp", "sampler.cpp", "ugs_sampler_batch_extension.cpp")]
            + [os.path.join(HERE, "ref_dump_shim.cpp")],
            os.path.join(ugs, "include"), v)
        eps = os.path.join(S, "epsilon_uniform_sampler")
        built[(v, "epsilon_uniform_sampler")] = _compile(
            "epsilon_uniform_sampler", [os.path.join(eps, "src", "epsilon_uniform_sampler.cpp")],
            os.path.join(eps, "include"), v, ["-fopenmp"], ["-fopenmp"])
        apx = os.path.join(S, "apx_ugs_sampler")
        built[(v, "apx_ugs_sampler")] = _compile(
            "apx_ugs_sampler", [os.path.join(apx, "src", "apx_ugs_sampler.cpp")],
            os.path.join(apx, "include"), v, ["-fopenmp"], ["-fopenmp"])
        uni = os.path.join(S, "uniform_sampler")
        built[(v, "uniform_sampler")] = _compile(
            "uniform_sampler", [os.path.join(uni, "src", "uniform_sampler.cpp")],
            os.path.join(uni, "include"), v, ["-fopenmp"], ["-fopenmp"])
        rwr = os.path.join(S, "rwr_sampler")
        built[(v, "rwr_sampler")] = _compile(
            "rwr_sampler", [os.path.join(rwr, "src", "rwr_sampler.cpp")],
            os.path.join(rwr, "src"), v, ["-fopenmp"], ["-fopenmp"])
    return built


if __name__ == "__main__":
    for k, p in build_all().items():
        print(k, p)
