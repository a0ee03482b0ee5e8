// op.cu -- the whole operator of the hot path behind ONE C-ABI call (include/lowbit_fa.h: lowbit_fa_fwd).
//
// Replaces the host orchestration of src/core.py:194-352 (lowbit_fa_qk_int8_pv_fp16_triton and its INT4 sibling):
//   km = k.mean(seq)                        core.py:293        -> lowbit_k_mean
//   q_int8, q_scale, k_int8, k_scale        core.py:300-319    -> lowbit_quant_per_block x 2 (K smoothing fused)
//   o (, lse) = attention forward           core.py:321-341    -> lowbit_attn_fwd
//   lse = lse / log2e + (q . km) sm_scale   core.py:343-350    -> lowbit_lse_fixup
// It launches exactly the kernels the separate entry points launch, with the same arguments, so the result is bit for bit
// the one of the multi-call path; what it removes is the host side of five ctypes calls and eight tensor allocations
// (measured at BASELINE config 1, B1 H2 N512 D64: 178 us of Python per call around 18 us of device work).
// The Q quantizer does not depend on the K chain (mean -> K codes): for tensors large enough to matter it runs on a
// per-device side stream forked from and joined to the caller's stream (capturable in a CUDA graph).
#include <mutex>

#include "common.cuh"

namespace lowbit {

struct OpLayout {
  int64_t q_codes, k_codes, q_scale, k_scale, km, mean_ws, total;
};
static inline int64_t align_up(int64_t v) { return (v + 255) & ~int64_t(255); }

static OpLayout op_layout(int B, int Hq, int Hkv, int Nq, int Nk, int D, int k_bits, int k_pack) {
  OpLayout L;
  const int64_t kd = (k_pack && k_bits < 8) ? (int64_t)D * k_bits / 8 : D;
  int64_t off = 0;
  L.q_codes = off; off = align_up(off + (int64_t)B * Hq * Nq * D);
  L.k_codes = off; off = align_up(off + (int64_t)B * Hkv * Nk * kd);
  L.q_scale = off; off = align_up(off + (int64_t)B * Hq * ((Nq + 127) / 128) * 4);
  L.k_scale = off; off = align_up(off + (int64_t)B * Hkv * ((Nk + 63) / 64) * 4);
  L.km = off; off = align_up(off + (int64_t)B * Hkv * D * 2);
  L.mean_ws = off; off = align_up(off + lowbit_k_mean_workspace_bytes(B, Hkv, Nk, D));
  L.total = off;
  return L;
}

// one side stream + fork / join events per device, created on first use
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream g_side[64];
static std::mutex g_side_mu;

static int side_stream(SideStream** out) {
  int dev = 0;
  LOWBIT_CUDA(cudaGetDevice(&dev));
  LOWBIT_CHECK(dev >= 0 && dev < 64, "lowbit_fa_fwd: device ordinal %d out of range", dev);
  SideStream& s = g_side[dev];
  if (s.st == nullptr) {
    LOWBIT_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    LOWBIT_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    LOWBIT_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return 0;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" int64_t lowbit_fa_fwd_workspace_bytes(int B, int Hq, int Hkv, int Nq, int Nk, int D, int k_bits, int k_pack) {
  return op_layout(B, Hq, Hkv, Nq, Nk, D, k_bits, k_pack).total;
}

extern "C" int lowbit_fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, void* workspace,
                             int B, int Hq, int Hkv, int Nq, int Nk, int D, int layout_nhd,
                             int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_n,
                             int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_n,
                             int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_n,
                             int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_n,
                             float sm_scale, float q_multiplier, int k_bits, int k_pack, int smooth_k, int quant_mode, int dtype,
                             int out_dtype, int flags, void* stream) {
  LOWBIT_CHECK(q && k && v && o && workspace, "lowbit_fa_fwd: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_fa_fwd: head_dim must be 64 or 128 (got %d; the host pads smaller ones)", D);
  LOWBIT_CHECK(k_bits == 8 || k_bits == 4, "lowbit_fa_fwd: k_bits must be 8 or 4 (got %d)", k_bits);
  LOWBIT_CHECK(Hkv > 0 && Hq % Hkv == 0, "lowbit_fa_fwd: Hq must be a multiple of Hkv");
  LOWBIT_CHECK(((uintptr_t)workspace & 255) == 0, "lowbit_fa_fwd: the workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const OpLayout L = op_layout(B, Hq, Hkv, Nq, Nk, D, k_bits, k_pack);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  void* q_codes = ws + L.q_codes;
  void* k_codes = ws + L.k_codes;
  float* q_scale = reinterpret_cast<float*>(ws + L.q_scale);
  float* k_scale = reinterpret_cast<float*>(ws + L.k_scale);
  void* km = smooth_k ? ws + L.km : nullptr;
  const int kd = (k_pack && k_bits < 8) ? D * k_bits / 8 : D;
  // codes live in the layout of their source tensor (contiguous)
  auto strides = [&](int H, int Nn, int dd, int64_t& sb, int64_t& sh, int64_t& sn) {
    if (layout_nhd) { sb = (int64_t)Nn * H * dd; sn = (int64_t)H * dd; sh = dd; }
    else { sb = (int64_t)H * Nn * dd; sh = (int64_t)Nn * dd; sn = dd; }
  };
  int64_t qcb, qch, qcn, kcb, kch, kcn;
  strides(Hq, Nq, D, qcb, qch, qcn);
  strides(Hkv, Nk, kd, kcb, kch, kcn);
  const float sm_q = q_multiplier;  // sm_scale * log2(e) as the caller rounds it (quant_per_block.py:205: the Q multiplier)

  // Q codes next to the K chain when the tensors are big enough for the overlap to pay (the fork / join costs ~5 us)
  const bool overlap = (int64_t)B * Hq * Nq * D >= (int64_t)1 << 22;
  SideStream* side = nullptr;
  cudaStream_t q_st = st;
  if (overlap) {
    std::lock_guard<std::mutex> g(g_side_mu);
    if (side_stream(&side) != 0) return 1;
    LOWBIT_CUDA(cudaEventRecord(side->fork, st));
    LOWBIT_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    q_st = side->st;
    if (lowbit_quant_per_block(q, nullptr, q_codes, q_scale, B, Hq, Nq, D, q_stride_b, q_stride_h, q_stride_n, qcb, qch,
                               qcn, 128, 8, 0, sm_q, quant_mode, dtype, q_st) != 0) return 1;
    LOWBIT_CUDA(cudaEventRecord(side->join, side->st));
  }
  if (smooth_k &&
      lowbit_k_mean(k, km, ws + L.mean_ws, B, Hkv, Nk, D, k_stride_b, k_stride_h, k_stride_n, dtype, st) != 0) return 1;
  if (lowbit_quant_per_block(k, km, k_codes, k_scale, B, Hkv, Nk, D, k_stride_b, k_stride_h, k_stride_n, kcb, kch, kcn,
                             64, k_bits, k_pack, 1.0f, quant_mode, dtype, st) != 0) return 1;
  if (overlap) {
    std::lock_guard<std::mutex> g(g_side_mu);
    LOWBIT_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  } else if (lowbit_quant_per_block(q, nullptr, q_codes, q_scale, B, Hq, Nq, D, q_stride_b, q_stride_h, q_stride_n, qcb,
                                    qch, qcn, 128, 8, 0, sm_q, quant_mode, dtype, st) != 0) {
    return 1;
  }
  const int qk_mode = (k_bits == 4 && k_pack) ? LOWBIT_QK_Q8K4 : LOWBIT_QK_I8;
  if (lowbit_attn_fwd(q_codes, k_codes, v, q_scale, k_scale, nullptr, nullptr, nullptr, o, lse, B, Hq, Hkv, Nq, Nk, D, qcb,
                      qch, qcn, kcb, kch, kcn, v_stride_b, v_stride_h, v_stride_n, o_stride_b, o_stride_h, o_stride_n,
                      qk_mode, LOWBIT_PV_F16, out_dtype, flags, st) != 0) return 1;
  if (lse != nullptr &&
      lowbit_lse_fixup(lse, q, km, B, Hq, Hkv, Nq, D, q_stride_b, q_stride_h, q_stride_n, sm_scale, dtype, st) != 0) return 1;
  return 0;
}
