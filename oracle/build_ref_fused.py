"""oracle/build_ref_fused.py -- TEST INFRASTRUCTURE: builds the reference's own CUDA quantizers as a checker.

Compiles, from the sources where they lie under /root/reference (nothing is copied into this repository):

    csrc/fused/fused.cu     QuantInt8Kernel :64-198, SubMeanKernel :200-261, TransposePadPermuteKernel :263-330,
                            MeanScaleKernel :332-428, launchers :430-980
    csrc/fused/pybind.cpp   the 8 entry points src/quant.py calls as `_fused.*` (src/quant.py:90-97,165-171,207,276-291)

with torch.utils.cpp_extension (the sources include <torch/extension.h>; the reference's own setup.py cannot run:
undefined CUDA_HOME at setup.py:42 and a Paddle build fed torch headers) into the git-ignored oracle/_ref/:

    oracle/_ref/fused_ieee/lowbit_ref_fused_ieee.so    plain nvcc -O3: IEEE division / multiplication -- the contract
                                                       the Q2 / Q6 rows are pinned to
    oracle/_ref/fused_fast/lowbit_ref_fused_fast.so    with the reference's own --use_fast_math (setup.py:66) -- kept to
                                                       document the delta (approximate division flips a few codes)

nvcc cross-compiles for sm_100 without a GPU, so this runs in the build container; the .so files travel to the GPU box
with the snapshot, where tools/make_golden_fused.py imports them (load()) and dumps tests/golden/fused_*.npz.
Only tests/ and tools/make_golden_fused.py use this module; the product never does.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/csrc/fused"
OUT = os.path.join(HERE, "_ref")
VARIANTS = {"ieee": [], "fast": ["--use_fast_math"]}


def so_path(variant):
    return os.path.join(OUT, f"fused_{variant}", f"lowbit_ref_fused_{variant}.so")


def build(variant="ieee", verbose=False):
    """Compile one variant (needs /root/reference: build container only)."""
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    d = os.path.join(OUT, f"fused_{variant}")
    os.makedirs(d, exist_ok=True)
    cpp_extension.load(
        name=f"lowbit_ref_fused_{variant}",
        sources=[os.path.join(REF, "fused.cu"), os.path.join(REF, "pybind.cpp")],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda"] + VARIANTS[variant],
        build_directory=d, verbose=verbose, is_python_module=False)
    return so_path(variant)


def load(variant="ieee"):
    """Import a prebuilt variant (GPU box: no /root/reference there, only the .so that travelled)."""
    import torch  # noqa: F401  (the extension links against libtorch)
    p = so_path(variant)
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} is missing: run `python oracle/build_ref_fused.py` in the build container")
    name = f"lowbit_ref_fused_{variant}"
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("reference sources not mounted; nothing to build", file=sys.stderr)
        sys.exit(0)
    for v in (sys.argv[1:] or list(VARIANTS)):
        print("built", build(v, verbose=False))
