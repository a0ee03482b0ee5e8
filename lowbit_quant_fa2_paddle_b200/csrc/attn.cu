// attn.cu -- fused low-bit FlashAttention forward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   _attn_fwd / _attn_fwd_inner   src/triton/attn_qk_int8_per_block.py:24-167        (non-causal)
//   _attn_fwd_base                src/triton/attn_qk_int8_per_block_causal.py:216-334 (causal)
//
// Per CTA: one 128-row Q tile of one (batch, q-head).  Q (int8), K (int8) and V (fp16) tiles are staged in
// shared memory by TMA (hardware swizzle, OOB rows zero-filled = the reference's masked loads);
// S = Q.K^T runs on tcgen05 kind::i8 (exact int32 in TMEM); each of the 128 threads owns one row (one TMEM
// lane): integer row-max, dequant by q_scale*k_scale, exp2 online softmax in registers, P written back to TMEM
// as fp16 (aliasing S) and consumed as the A operand of the P.V tcgen05 kind::f16 MMA whose fp32
// accumulator O stays resident in TMEM for the whole key loop (rescaled only when the running max moves by
// more than 2^8).  Epilogue: O/l -> fp16/bf16, lse2 = log2(l) + m.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <limits.h>

namespace lowbit {

struct AttnParams {
  const float* q_scale;
  const float* k_scale;
  void* o;
  float* lse;
  int Hq, Hkv, Nq, Nk;
  int nqb, nkb;  // scale blocks per (b,h): ceil(Nq/128), ceil(Nk/64)
  int64_t osb, osh, osn;
  int flags, out_dtype;
  int32_t* dbg;  // diagnostics: when non-null, CTA (0,0,0) dumps the int32 scores of key block 0 ([128][64])
};

static int32_t* g_attn_debug = nullptr;

constexpr int kBM = 128;  // Q rows per CTA (= TMEM lanes)
constexpr int kBN = 64;   // keys per step (= the reference's k_scale granularity)

template <int D>
struct AttnSmem {
  static constexpr int kQ = kBM * D;       // int8
  static constexpr int kK = kBN * D;       // int8
  static constexpr int kV = kBN * D * 2;   // fp16
  static constexpr int kStages = 2;
  static constexpr int kBytes = kQ + kStages * (kK + kV) + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using SM = AttnSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SM::kQ;
  uint8_t* sV = sK + SM::kStages * SM::kK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + SM::kStages * SM::kV);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv = bars + 1;  // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heavy causal tiles first
  const int hq = blockIdx.y, b = blockIdx.z;
  const int hkv = hq / (p.Hq / p.Hkv);

  constexpr uint32_t kTmemCols = (D == 64) ? 128 : 256;  // S/P: [0,64)  O: [64, 64+D)
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  if (tid == 32) {
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_kv + 0, 1);
    ptx::mbar_init(bar_kv + 1, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;         // int32 scores, 64 columns
  const uint32_t tP = tmem_base;         // fp16 probabilities alias S (32 columns)
  const uint32_t tO = tmem_base + kBN;   // fp32 output accumulator, D columns
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  // key-block range of this Q tile
  int nblk = p.nkb;
  if (CAUSAL) nblk = min(nblk, (qt + 1) * (kBM / kBN));

  constexpr uint32_t kKVBytes = SM::kK + SM::kV;
  auto load_kv = [&](int j, int st) {
    ptx::mbar_expect_tx(bar_kv + st, kKVBytes);
    ptx::tma_load_4d(sK + st * SM::kK, &tmK, bar_kv + st, 0, j * kBN, hkv, b);
    ptx::tma_load_4d(sV + st * SM::kV, &tmV, bar_kv + st, 0, j * kBN, hkv, b);
    if (D == 128) ptx::tma_load_4d(sV + st * SM::kV + kBN * 128, &tmV, bar_kv + st, 64, j * kBN, hkv, b);
  };
  if (tid == 0) {
    ptx::mbar_expect_tx(bar_q, SM::kQ);
    ptx::tma_load_4d(sQ, &tmQ, bar_q, 0, qt * kBM, hq, b);
    load_kv(0, 0);
  }

  // descriptors
  constexpr uint32_t kSwzQK = (D == 64) ? ptx::kSwz64 : ptx::kSwz128;
  constexpr uint32_t kSboQK = 8 * D;  // 8 rows of D bytes
  constexpr uint32_t idesc_qk = ptx::make_idesc(ptx::kCS32, ptx::kS8, ptx::kS8, 0, 0, kBM, kBN);
  constexpr uint32_t idesc_pv = ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 1, kBM, D);

  const int row = qt * kBM + tid;  // global query row owned by this thread
  const float qs = p.q_scale[((int64_t)b * p.Hq + hq) * p.nqb + qt];
  const float* ks_ptr = p.k_scale + ((int64_t)b * p.Hkv + hkv) * p.nkb;
  const bool mask_tail = !(p.flags & LOWBIT_ATTN_COMPAT_TAIL) && (p.Nk % kBN != 0);

  float m_ref = -INFINITY, l = 0.f;

  for (int j = 0; j < nblk; ++j) {
    const int st = j & 1;
    const uint32_t ph = j & 1;
    if (tid == 0) {
      if (j + 1 < nblk) load_kv(j + 1, st ^ 1);  // stage st^1 was released by the PV MMA of step j-1
      if (j == 0) ptx::mbar_wait(bar_q, 0, 1);
      ptx::mbar_wait(bar_kv + st, (j >> 1) & 1, 2);
      ptx::tc_fence_after();
      const uint32_t aq = ptx::smem_u32(sQ), ak = ptx::smem_u32(sK + st * SM::kK);
#pragma unroll
      for (int kk = 0; kk < D / 32; ++kk) {
        const uint64_t da = ptx::make_smem_desc(aq + kk * 32, 16, kSboQK, kSwzQK);
        const uint64_t db = ptx::make_smem_desc(ak + kk * 32, 16, kSboQK, kSwzQK);
        ptx::umma_i8_ss(tS, da, db, idesc_qk, kk > 0);
      }
      ptx::umma_commit(bar_s);
    }
    ptx::mbar_wait(bar_s, ph, 3);
    ptx::tc_fence_after();

    // ---- softmax for this thread's row -------------------------------------------------------
    uint32_t s[kBN];
    ptx::tmem_ld_x32(tS + lane_off, s);
    ptx::tmem_ld_x32(tS + lane_off + 32, s + 32);
    ptx::tmem_wait_ld();

    if (p.dbg != nullptr && j == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
      for (int c = 0; c < kBN; ++c) p.dbg[tid * kBN + c] = (int)s[c];
    }
    const int c0 = j * kBN;
    int lim = kBN;  // columns [0, lim] are live
    if (CAUSAL && c0 + kBN - 1 > qt * kBM) lim = min(lim, row - c0);
    if (mask_tail && j == p.nkb - 1) lim = min(lim, p.Nk - 1 - c0);
    const bool masked = lim < kBN - 1;

    int imax = INT_MIN;
    if (!masked) {
#pragma unroll
      for (int c = 0; c < kBN; ++c) imax = max(imax, (int)s[c]);
    } else {
#pragma unroll
      for (int c = 0; c < kBN; ++c) imax = max(imax, c <= lim ? (int)s[c] : INT_MIN);
    }
    const float sc = qs * ks_ptr[j];
    const float mblk = (imax == INT_MIN) ? -INFINITY : (float)imax * sc;
    // lazy rescale: move the reference max only when it grows by more than 2^8 (warp-uniform decision,
    // tcgen05.ld/st are warp collectives)
    const bool need = mblk > m_ref + 8.f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = fmaxf(m_ref, mblk);
      const float alpha = (m_new == -INFINITY) ? 1.f : ptx::ex2(m_ref - m_new);  // m_ref == -inf -> 0
      l *= alpha;
      m_ref = m_new;
      if (j > 0) {
#pragma unroll
        for (int c = 0; c < D; c += 32) {
          uint32_t o[32];
          ptx::tmem_ld_x32(tO + lane_off + c, o);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          ptx::tmem_st_x32(tO + lane_off + c, o);
        }
      }
    }
    uint32_t pk[kBN / 2];
    float lsum = 0.f;
    const float neg_m = -m_ref;
#pragma unroll
    for (int c = 0; c < kBN; c += 2) {
      float p0 = ptx::ex2(fmaf(ptx::i2f_small((int)s[c]), sc, neg_m));
      float p1 = ptx::ex2(fmaf(ptx::i2f_small((int)s[c + 1]), sc, neg_m));
      if (masked) {
        p0 = (c <= lim) ? p0 : 0.f;
        p1 = (c + 1 <= lim) ? p1 : 0.f;
      }
      lsum += p0 + p1;
      pk[c / 2] = ptx::pack_f16x2(p0, p1);
    }
    l += lsum;
    ptx::tmem_st_x32(tP + lane_off, pk);
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t av = ptx::smem_u32(sV + st * SM::kV);
#pragma unroll
      for (int kk = 0; kk < kBN / 16; ++kk) {
        // V tile: MN-major (d contiguous), 128B swizzle: 8 key rows = 1024 B (SBO); 64-wide d atoms kBN*128 B apart (LBO)
        const uint64_t db = ptx::make_smem_desc(av + kk * 16 * 128, kBN * 128, 1024, ptx::kSwz128);
        ptx::umma_f16_ts(tO, tP + kk * 8, db, idesc_pv, (j > 0) || (kk > 0));
      }
      ptx::umma_commit(bar_o);
    }
    ptx::mbar_wait(bar_o, ph, 4);
    ptx::tc_fence_after();
  }

  // ---- epilogue: O / l -> out dtype, lse2 = log2(l) + m --------------------------------------------
  const float inv_l = 1.0f / l;
  const bool live_row = row < p.Nq;
  uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) + ((int64_t)b * p.osb + (int64_t)hq * p.osh + (int64_t)row * p.osn) * 2;
#pragma unroll
  for (int c = 0; c < D; c += 32) {
    uint32_t o[32];
    ptx::tmem_ld_x32(tO + lane_off + c, o);  // warp collective: every lane executes it
    ptx::tmem_wait_ld();
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = __uint_as_float(o[2 * i]) * inv_l, bb = __uint_as_float(o[2 * i + 1]) * inv_l;
      w[i] = (p.out_dtype == LOWBIT_F16) ? ptx::pack_f16x2(a, bb) : ptx::pack_bf16x2(a, bb);
    }
    if (live_row) {
      uint4* dst = reinterpret_cast<uint4*>(orow + c * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
  if (live_row && p.lse) p.lse[((int64_t)b * p.Hq + hq) * p.Nq + row] = ptx::lg2(l) + m_ref;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// logical [B,H,N,D] tensor with element strides -> 4-D tensor map (d, n, h, b), box (box_d, box_n, 1, 1)
static int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, int B, int H, int N, int D,
                    int64_t sb, int64_t sh, int64_t sn, int box_d, int box_n, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  LOWBIT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  LOWBIT_CHECK(((uintptr_t)ptr & 15) == 0, "tensor base address must be 16-byte aligned");
  LOWBIT_CHECK((sn * esize) % 16 == 0 && (sh * esize) % 16 == 0 && (sb * esize) % 16 == 0,
               "tensor strides must be multiples of 16 bytes");
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(sn * esize), (cuuint64_t)(sh * esize), (cuuint64_t)(sb * esize)};
  // a size-1 dimension may carry any stride in the caller's tensor; TMA still wants a legal (non-zero, 16B) one
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = 16;
  cuuint32_t box[4] = {(cuuint32_t)box_d, (cuuint32_t)box_n, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOWBIT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

template <int D, bool CAUSAL>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                       cudaStream_t st) {
  auto kern = attn_fwd_kernel<D, CAUSAL>;
  static bool configured = false;
  if (!configured) {
    LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<D>::kBytes));
    configured = true;
  }
  dim3 grid((p.Nq + kBM - 1) / kBM, p.Hq, B);
  kern<<<grid, 128, AttnSmem<D>::kBytes, st>>>(tq, tk, tv, p);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" void lowbit_attn_set_debug_buffer(void* dev_buf) { lowbit::g_attn_debug = (int32_t*)dev_buf; }

extern "C" int lowbit_attn_fwd(const void* q_codes, const void* k_codes, const void* v, const float* q_scale,
                               const float* k_scale, const float* v_scale, const float* v_mean, const int32_t* kbits,
                               void* o, float* lse, int B, int Hq, int Hkv, int Nq, int Nk, int D,
                               int64_t qsb, int64_t qsh, int64_t qsn, int64_t ksb, int64_t ksh, int64_t ksn,
                               int64_t vsb, int64_t vsh, int64_t vsn, int64_t osb, int64_t osh, int64_t osn,
                               int qk_mode, int pv_mode, int out_dtype, int flags, void* stream) {
  LOWBIT_CHECK(q_codes && k_codes && v && q_scale && k_scale && o, "lowbit_attn_fwd: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_attn_fwd: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(B > 0 && Hq > 0 && Hkv > 0 && Nq > 0 && Nk > 0, "lowbit_attn_fwd: empty tensor");
  LOWBIT_CHECK(Hq % Hkv == 0, "lowbit_attn_fwd: num_qo_heads (%d) must be divisible by num_kv_heads (%d)", Hq, Hkv);
  LOWBIT_CHECK(out_dtype == LOWBIT_F16 || out_dtype == LOWBIT_BF16, "lowbit_attn_fwd: bad out_dtype %d", out_dtype);
  LOWBIT_CHECK(qk_mode == LOWBIT_QK_I8, "lowbit_attn_fwd: qk_mode %d not implemented yet", qk_mode);
  LOWBIT_CHECK(pv_mode == LOWBIT_PV_F16, "lowbit_attn_fwd: pv_mode %d not implemented yet", pv_mode);
  const bool causal = flags & LOWBIT_ATTN_CAUSAL;
  LOWBIT_CHECK(!causal || Nq == Nk, "lowbit_attn_fwd: causal attention requires qo_len == kv_len");
  LOWBIT_CHECK((osn % 8) == 0 && (osh % 8) == 0 && (osb % 8) == 0 && ((uintptr_t)o & 15) == 0,
               "lowbit_attn_fwd: output must keep 16-byte row alignment");
  (void)v_scale; (void)v_mean; (void)kbits;
  cudaStream_t st = (cudaStream_t)stream;

  CUtensorMap tq, tk, tv;
  const CUtensorMapSwizzle swz_qk = (D == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (make_map(&tq, q_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, Hq, Nq, D, qsb, qsh, qsn, D, kBM, swz_qk)) return 1;
  if (make_map(&tk, k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, Hkv, Nk, D, ksb, ksh, ksn, D, kBN, swz_qk)) return 1;
  if (make_map(&tv, v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, Hkv, Nk, D, vsb, vsh, vsn, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;

  AttnParams p;
  p.q_scale = q_scale; p.k_scale = k_scale; p.o = o; p.lse = lse;
  p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
  p.nqb = (Nq + 127) / 128; p.nkb = (Nk + 63) / 64;
  p.osb = osb; p.osh = osh; p.osn = osn;
  p.flags = flags; p.out_dtype = out_dtype; p.dbg = g_attn_debug;
  if (D == 64) return causal ? launch_attn<64, true>(tq, tk, tv, p, B, st) : launch_attn<64, false>(tq, tk, tv, p, B, st);
  return causal ? launch_attn<128, true>(tq, tk, tv, p, B, st) : launch_attn<128, false>(tq, tk, tv, p, B, st);
}
