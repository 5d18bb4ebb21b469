"""TEST INFRASTRUCTURE ONLY -- never imported by rick_b200/ or by bench.py's GPU arm.

Compiles the reference's OWN native sources for this path -- op/upfirdn2d.cpp + op/upfirdn2d_kernel.cu and
op/fused_bias_act.cpp + op/fused_bias_act_kernel.cu, read where they lie under /root/reference (nothing is copied into
this repository) -- for sm_100a into ``oracle/_ref/`` (git-ignored, travels to the GPU box with the snapshot):

    oracle/_ref/ref_upfirdn2d.so     Python extension module exporting ``upfirdn2d(input, kernel, up_x, ..., pad_y1)``
    oracle/_ref/ref_fused.so         Python extension module exporting ``fused_bias_act(input, bias, refer, act, grad, alpha, scale)``

These are the exact binaries the reference would JIT-build at import (op/upfirdn2d.py:10-16, op/fused_act.py:10-16).
On a B200 they serve as (1) a second oracle: the GPU parity tests compare rick_b200's kernels with the reference's own
CUDA kernels on the same inputs, and (2) the kernel-to-beat of SURVEY.md section 2a (tests/test_ref_cuda_gpu.py prints
both timings).  The recipe is two nvcc/g++ compile lines per module through torch.utils.cpp_extension's ninja driver;
the reference's own build system (JIT at import) is not run.

    python oracle/build_ref_cuda.py            # no-op when /root/reference is absent or the modules are up to date
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_OP = os.environ.get("RICK_REFERENCE_OP", "/root/reference/op")

MODULES = {
    "ref_upfirdn2d": ("upfirdn2d.cpp", "upfirdn2d_kernel.cu"),
    "ref_fused": ("fused_bias_act.cpp", "fused_bias_act_kernel.cu"),
}


def _stale(name: str, sources) -> bool:
    so = os.path.join(OUT, name + ".so")
    if not os.path.isfile(so):
        return True
    return any(os.path.getmtime(s) > os.path.getmtime(so) for s in sources)


def build(verbose: bool = False) -> bool:
    """Returns True when both modules exist afterwards; False (and builds nothing) when the reference is not here."""
    if not os.path.isdir(REF_OP):
        return all(os.path.isfile(os.path.join(OUT, n + ".so")) for n in MODULES)
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")     # no GPU here: keep torch from probing for one
    from torch.utils import cpp_extension
    for name, files in MODULES.items():
        sources = [os.path.join(REF_OP, f) for f in files]
        if not _stale(name, sources):
            continue
        build_dir = os.path.join(OUT, "build_" + name)
        os.makedirs(build_dir, exist_ok=True)
        cpp_extension.load(
            name=name, sources=sources, build_directory=build_dir, is_python_module=False, verbose=verbose,
            extra_cflags=["-O2"],
            extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        os.replace(os.path.join(build_dir, name + ".so"), os.path.join(OUT, name + ".so"))
    return True


def load(name: str):
    """Import oracle/_ref/<name>.so as a Python extension module (None when it has not been built)."""
    import importlib.machinery
    import importlib.util
    so = os.path.join(OUT, name + ".so")
    if not os.path.isfile(so):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    loader = importlib.machinery.ExtensionFileLoader(name, so)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference sources not present, nothing built")
