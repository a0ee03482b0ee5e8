"""oracle/build_ref_qattn.py -- TEST INFRASTRUCTURE: builds the reference's own CUDA attention kernels as a checker.

Compiles, from the sources where they lie under /root/reference (nothing is copied into this repository):

    csrc/qattn/qk_int_sv_f8_cuda.cu    INT8 Q.K^T + FP8 (e4m3) P.V, fp32 accumulation -- the A3 semantics
                                       (qk_int8_sv_f8_accum_f32_attn, ..._fuse_v_scale_attn, ..._fuse_v_scale_fuse_v_mean_attn)
    csrc/qattn/qk_int_sv_f16_cuda.cu   INT8 Q.K^T + FP16 P.V (qk_int8_sv_f16_accum_f32_attn, accum_f16 variants)

plus oracle/ref_qattn_bind.cpp, a pybind stub of OURS that binds the entry points csrc/qattn/attn_cuda.h declares (the
reference's own pybind.cpp also binds the *_buf kernels of three more translation units, which are not needed here).
The kernels are mma.sync code for sm80 / sm89 (no wgmma): they compile for sm_100 and run on a B200.  Output:
oracle/_ref/qattn/lowbit_ref_qattn.so (git-ignored, travels to the GPU box).  Used by tools/make_golden_qattn.py only.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/csrc/qattn"
OUT = os.path.join(HERE, "_ref", "qattn")
NAME = "lowbit_ref_qattn"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    os.makedirs(OUT, exist_ok=True)
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(REF, "qk_int_sv_f8_cuda.cu"), os.path.join(REF, "qk_int_sv_f16_cuda.cu"),
                 os.path.join(HERE, "ref_qattn_bind.cpp")],
        extra_include_paths=[REF],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                           "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
                           "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
                           "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
                           "--threads=4", "-Xptxas=-v", "-diag-suppress=174"],
        build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path()


def load():
    import torch  # noqa: F401
    p = so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} is missing: run `python oracle/build_ref_qattn.py` in the build container")
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("reference sources not mounted; nothing to build", file=sys.stderr)
        sys.exit(0)
    print("built", build(verbose="-v" in sys.argv))
