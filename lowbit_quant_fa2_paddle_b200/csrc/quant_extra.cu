// quant_extra.cu -- Q3 per-thread quantizer, Q5 KIVI pack, Q6 V->FP8 (placeholders until implemented)
#include "common.cuh"
using namespace lowbit;
extern "C" {
int lowbit_quant_per_thread(const void*, const void*, void*, float*, int, int, int, int, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int64_t, int, int, int, int, int, void*) {
  return fail("lowbit_quant_per_thread: not implemented yet");
}
int lowbit_quant_pack_lastdim(const void*, void*, void*, void*, int64_t, int, int, int, int, void*) {
  return fail("lowbit_quant_pack_lastdim: not implemented yet");
}
int lowbit_v_fp8_per_channel(const void*, void*, float*, float*, int, int, int, int, int64_t, int64_t, int64_t, float,
                             int, void*) {
  return fail("lowbit_v_fp8_per_channel: not implemented yet");
}
int lowbit_attn_fwd_partial(const void*, const void*, const void*, const float*, const float*, float*, float*, float*,
                            int, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int64_t, int64_t, int, int, int, void*) {
  return fail("lowbit_attn_fwd_partial: not implemented yet");
}
int lowbit_attn_finalize(const float*, const float*, const float*, void*, float*, int, int, int, int, int64_t, int64_t,
                         int64_t, int, void*) {
  return fail("lowbit_attn_finalize: not implemented yet");
}
}
