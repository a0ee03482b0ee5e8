// oracle/ref_qattn_bind.cpp -- TEST INFRASTRUCTURE.  pybind stub (ours) over the entry points the reference declares in
// csrc/qattn/attn_cuda.h and defines in qk_int_sv_f8_cuda.cu / qk_int_sv_f16_cuda.cu; built by oracle/build_ref_qattn.py.
#include <torch/extension.h>

#include "attn_cuda.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("qk_int8_sv_f16_accum_f32_attn", &qk_int8_sv_f16_accum_f32_attn);
  m.def("qk_int8_sv_f16_accum_f16_attn", &qk_int8_sv_f16_accum_f16_attn);
  m.def("qk_int8_sv_f16_accum_f16_fuse_v_mean_attn", &qk_int8_sv_f16_accum_f16_fuse_v_mean_attn);
  m.def("qk_int8_sv_f8_accum_f32_attn", &qk_int8_sv_f8_accum_f32_attn);
  m.def("qk_int8_sv_f8_accum_f32_fuse_v_scale_attn", &qk_int8_sv_f8_accum_f32_fuse_v_scale_attn);
  m.def("qk_int8_sv_f8_accum_f32_fuse_v_scale_fuse_v_mean_attn", &qk_int8_sv_f8_accum_f32_fuse_v_scale_fuse_v_mean_attn);
}
