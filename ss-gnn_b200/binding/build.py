"""Builds ugs_sampler_b200_binding.so: the pybind11/libtorch binding of the C ABI that INTEGRATION.md shows
(what a reference maintainer would add next to src/samplers/ugs_sampler/src/extension.cpp)."""
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
REPO = os.path.dirname(PKG)
SRC = os.path.join(HERE, "ugs_sampler_b200_binding.cpp")
OUT = os.path.join(HERE, "ugs_sampler_b200_binding.so")


def build(force: bool = False) -> str:
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce
    deps = [SRC, os.path.join(REPO, "include", "ugs_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    inc = ce.include_paths() + [pybind11.get_include(), sysconfig.get_paths()["include"], os.path.join(REPO, "include"), cuda_inc]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-DTORCH_EXTENSION_NAME=ugs_sampler_b200_binding",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + [f"-I{p}" for p in inc] + \
          [SRC, "-o", OUT, f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python",
           f"-Wl,-rpath,{tlib}", f"-L{os.path.join(PKG, 'ugs_b200')}", "-lugs_b200", "-Wl,-rpath,$ORIGIN/../ugs_b200"]
    subprocess.check_call(cmd)
    return OUT


def load():
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401
    path = build()
    loader = importlib.machinery.ExtensionFileLoader("ugs_sampler_b200_binding", path)
    spec = importlib.util.spec_from_file_location("ugs_sampler_b200_binding", path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod
