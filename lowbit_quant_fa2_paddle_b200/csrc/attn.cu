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
#include <stdlib.h>
#include <type_traits>

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

constexpr int kBM = 128;      // Q rows per CTA (= TMEM lanes)
constexpr int kScaleBlk = 64; // k_scale granularity of the reference quantizer (BLKK)
constexpr int kSoftmaxThreads = 128;
constexpr int kThreads = kSoftmaxThreads + 32;  // + one helper warp (TMA producer + tcgen05 issuer, one elected lane)

// Per-head-dim tiling.  D=64 is exp2(MUFU)-bound: small 32-key steps keep the register footprint under 96 so that
// four CTAs (16 softmax warps) share an SM and hide each other's latencies.  D=128 has twice the tensor work per
// exp2 and runs 64-key steps with two CTAs per SM.
template <int D> struct AttnCfg;
template <> struct AttnCfg<64> {
  static constexpr int BN = 32, CTAS = 4, KS = 4, VS = 3, TMEM_COLS = 128;
};
template <> struct AttnCfg<128> {
  static constexpr int BN = 64, CTAS = 2, KS = 4, VS = 3, TMEM_COLS = 256;
};

template <int D>
struct AttnSmem {
  using C = AttnCfg<D>;
  static constexpr int kQ = kBM * D;        // int8
  static constexpr int kK = C::BN * D;      // int8
  static constexpr int kV = C::BN * D * 2;  // fp16
  static constexpr int kBytes = kQ + C::KS * kK + C::VS * kV + 256 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int VAR>
__device__ __forceinline__ float score_to_f32(uint32_t v) {
  return (VAR & 1) ? ptx::i2f_small((int)v) : __int2float_rn((int)v);
}
// One softmax step over a BN-key block for one query row: p = exp2(S*sc - m), packed to fp16 pairs, row sum in fp32.
// MASKED: columns > lim contribute 0 (causal diagonal band / masked tail keys).
template <int BN, bool MASKED, int VAR>
__device__ __forceinline__ float softmax_block(const uint32_t (&s)[BN], float sc, float neg_m, int lim,
                                               uint32_t (&pk)[BN / 2]) {
  if constexpr ((VAR & 2) == 0) {
    // packed fp32x2 arithmetic (FFMA2 / FADD2): one instruction scales, or accumulates, two scores
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < BN; c += 4) {
      const float2 x0 = __ffma2_rn(make_float2(score_to_f32<VAR>(s[c]), score_to_f32<VAR>(s[c + 1])), sc2, nm2);
      const float2 x1 = __ffma2_rn(make_float2(score_to_f32<VAR>(s[c + 2]), score_to_f32<VAR>(s[c + 3])), sc2, nm2);
      float2 p0 = make_float2(ptx::ex2(x0.x), ptx::ex2(x0.y));
      float2 p1 = make_float2(ptx::ex2(x1.x), ptx::ex2(x1.y));
      if (MASKED) {
        p0.x = (c <= lim) ? p0.x : 0.f;
        p0.y = (c + 1 <= lim) ? p0.y : 0.f;
        p1.x = (c + 2 <= lim) ? p1.x : 0.f;
        p1.y = (c + 3 <= lim) ? p1.y : 0.f;
      }
      acc0 = __fadd2_rn(acc0, p0);
      acc1 = __fadd2_rn(acc1, p1);
      pk[c / 2] = ptx::pack_f16x2(p0.x, p0.y);
      pk[c / 2 + 1] = ptx::pack_f16x2(p1.x, p1.y);
    }
    const float2 t = __fadd2_rn(acc0, acc1);
    return t.x + t.y;
  } else {
    float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
    for (int c = 0; c < BN; c += 2) {
      float p0 = ptx::ex2(fmaf(score_to_f32<VAR>(s[c]), sc, neg_m));
      float p1 = ptx::ex2(fmaf(score_to_f32<VAR>(s[c + 1]), sc, neg_m));
      if (MASKED) {
        p0 = (c <= lim) ? p0 : 0.f;
        p1 = (c + 1 <= lim) ? p1 : 0.f;
      }
      lsum0 += p0;
      lsum1 += p1;
      pk[c / 2] = ptx::pack_f16x2(p0, p1);
    }
    return lsum0 + lsum1;
  }
}
template <int BN, bool MASKED>
__device__ __forceinline__ int row_max(const uint32_t (&s)[BN], int lim) {
  int m4[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};  // 4 independent chains (latency, not throughput, bound)
#pragma unroll
  for (int c = 0; c < BN; ++c) m4[c & 3] = max(m4[c & 3], (!MASKED || c <= lim) ? (int)s[c] : INT_MIN);
  return max(max(m4[0], m4[1]), max(m4[2], m4[3]));
}
template <int N> __device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 32) ptx::tmem_ld_x32(taddr, r);
  else { ptx::tmem_ld_x32(taddr, r); ptx::tmem_ld_x32(taddr + 32, r + 32); }
}
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
  if constexpr (N == 16) ptx::tmem_st_x16(taddr, r);
  else ptx::tmem_st_x32(taddr, r);
}

// Warp roles:  warps 0-3  softmax (thread t <-> query row t <-> TMEM lane t)
//              warp 4     helper: one elected lane is both the TMA producer and the tcgen05 issuer; the warp also
//                         owns the TMEM allocation
// TMEM columns: S/P buffer 0 [0,BN)  S/P buffer 1 [BN,2BN)  O [2BN, 2BN+D)
// Pipeline: QK_{j+2} is issued right after PV_j, so the int8 contraction of the next two key blocks and the fp16
// P.V of the previous one run on the tensor pipe while the softmax warps work on block j; K/V stages are refilled
// by the same thread as soon as the MMAs that read them have committed.
template <int D, bool CAUSAL, int VAR, bool DBG>
__global__ void __launch_bounds__(kThreads, AttnCfg<D>::CTAS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<D>;
  using SM = AttnSmem<D>;
  constexpr int BN = C::BN, KS = C::KS, VS = C::VS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SM::kQ;
  uint8_t* sV = sK + KS * SM::kK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + VS * SM::kV);
  uint64_t* bar_q = bars + 0;
  uint64_t* kfull = bars + 1;        // [KS] TMA -> MMA
  uint64_t* kfree = kfull + KS;      // [KS] MMA (commit) -> TMA
  uint64_t* vfull = kfree + KS;      // [VS]
  uint64_t* vfree = vfull + VS;      // [VS]
  uint64_t* bar_s = vfree + VS;      // [2] QK done: S buffer b holds scores
  uint64_t* p_ready = bar_s + 2;     // [2] 128 softmax threads wrote P into buffer b
  uint64_t* bar_o = p_ready + 2;     // PV_j done (one phase per key block)
  uint64_t* bar_final = bar_o + 1;   // last PV done (single phase: parity waits must never lag 2 phases)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_final + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heavy causal tiles first
  const int hq = blockIdx.y, b = blockIdx.z;
  const int hkv = hq / (p.Hq / p.Hkv);

  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (tid == 0) {
    ptx::mbar_init(bar_q, 1);
    for (int i = 0; i < KS; ++i) { ptx::mbar_init(kfull + i, 1); ptx::mbar_init(kfree + i, 1); }
    for (int i = 0; i < VS; ++i) { ptx::mbar_init(vfull + i, 1); ptx::mbar_init(vfree + i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_s + i, 1); ptx::mbar_init(p_ready + i, kSoftmaxThreads); }
    ptx::mbar_init(bar_o, 1);
    ptx::mbar_init(bar_final, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 2 * BN;  // fp32 output accumulator, D columns

  // key-block range of this Q tile.  compat_tail walks the reference's whole 64-key blocks (phantom zero keys).
  const bool compat = (p.flags & LOWBIT_ATTN_COMPAT_TAIL) != 0;
  const int nk_eff = compat ? p.nkb * kScaleBlk : p.Nk;
  int nblk = (nk_eff + BN - 1) / BN;
  if (CAUSAL) nblk = min(nblk, (qt + 1) * (kBM / BN));

  if (warp == 4) {
    // ================================ helper: TMA producer + tcgen05 issuer ================================
    if (ptx::elect_one()) {
      constexpr uint32_t kSwzQK = (D == 64) ? ptx::kSwz64 : ptx::kSwz128;
      constexpr uint32_t kSboQK = 8 * D;  // 8 rows of D bytes
      constexpr uint32_t idesc_qk = ptx::make_idesc(ptx::kCS32, ptx::kS8, ptx::kS8, 0, 0, kBM, BN);
      constexpr uint32_t idesc_pv = ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 1, kBM, D);
      const uint32_t aq = ptx::smem_u32(sQ);
      auto load_k = [&](int j) {
        const int ks = j % KS;
        ptx::mbar_wait(kfree + ks, ((j / KS) & 1) ^ 1, 10);
        ptx::mbar_expect_tx(kfull + ks, SM::kK);
        ptx::tma_load_4d(sK + ks * SM::kK, &tmK, kfull + ks, 0, j * BN, hkv, b);
      };
      auto load_v = [&](int j) {
        const int vs = j % VS;
        ptx::mbar_wait(vfree + vs, ((j / VS) & 1) ^ 1, 11);
        ptx::mbar_expect_tx(vfull + vs, SM::kV);
        ptx::tma_load_4d(sV + vs * SM::kV, &tmV, vfull + vs, 0, j * BN, hkv, b);
        if (D == 128) ptx::tma_load_4d(sV + vs * SM::kV + BN * 128, &tmV, vfull + vs, 64, j * BN, hkv, b);
      };
      auto issue_qk = [&](int j) {
        const int ks = j % KS;
        ptx::mbar_wait(kfull + ks, (j / KS) & 1, 20);
        ptx::tc_fence_after();
        const uint32_t ak = ptx::smem_u32(sK + ks * SM::kK);
        const uint32_t tS = tmem_base + (j & 1) * BN;
#pragma unroll
        for (int kk = 0; kk < D / 32; ++kk) {
          const uint64_t da = ptx::make_smem_desc(aq + kk * 32, 16, kSboQK, kSwzQK);
          const uint64_t db = ptx::make_smem_desc(ak + kk * 32, 16, kSboQK, kSwzQK);
          ptx::umma_i8_ss(tS, da, db, idesc_qk, kk > 0);
        }
        ptx::umma_commit(bar_s + (j & 1));  // scores ready for the softmax warps
        ptx::umma_commit(kfree + ks);       // K stage may be refilled
      };
      ptx::mbar_expect_tx(bar_q, SM::kQ);
      ptx::tma_load_4d(sQ, &tmQ, bar_q, 0, qt * kBM, hq, b);
      for (int j = 0; j < min(KS, nblk); ++j) load_k(j);
      for (int j = 0; j < min(2, nblk); ++j) load_v(j);
      ptx::mbar_wait(bar_q, 0, 21);
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        const int vs = j % VS;
        ptx::mbar_wait(p_ready + (j & 1), (j >> 1) & 1, 22);
        ptx::mbar_wait(vfull + vs, (j / VS) & 1, 23);
        ptx::tc_fence_after();
        const uint32_t av = ptx::smem_u32(sV + vs * SM::kV);
        const uint32_t tP = tmem_base + (j & 1) * BN;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          // V tile: MN-major (d contiguous), 128B swizzle: 8 key rows = 1024 B (SBO); 64-wide d atoms BN*128 B apart (LBO)
          const uint64_t db = ptx::make_smem_desc(av + kk * 16 * 128, BN * 128, 1024, ptx::kSwz128);
          ptx::umma_f16_ts(tO, tP + kk * 8, db, idesc_pv, (j > 0) || (kk > 0));
        }
        ptx::umma_commit(vfree + vs);
        ptx::umma_commit(bar_o);
        if (j == nblk - 1) ptx::umma_commit(bar_final);
        if (j + 2 < nblk) issue_qk(j + 2);  // overwrites S/P buffer (j&1): ordered after PV_j on the tensor pipe
        // refill: K stage of QK_j (long complete) and V stage of PV_{j-1} (complete in steady state)
        if (j + KS < nblk) load_k(j + KS);
        if (j + 2 < nblk) load_v(j + 2);
      }
    }
  } else {
    // ================================ softmax warps ================================
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int row = qt * kBM + tid;  // global query row owned by this thread
    const float qs = p.q_scale[((int64_t)b * p.Hq + hq) * p.nqb + qt];
    const float* ks_ptr = p.k_scale + ((int64_t)b * p.Hkv + hkv) * p.nkb;
    const bool mask_tail = !compat && (p.Nk % BN != 0);
    const int last_kblk = (p.Nk + BN - 1) / BN - 1;
    const uint32_t tS0 = tmem_base + lane_off, tS1 = tS0 + BN, tOl = tO + lane_off;
    float m_ref = -INFINITY, l = 0.f;

    // one key block: wait for S, row max, (rare) rescale of O, P = exp2(S*sc - m) -> TMEM, signal the issuer
    auto step = [&](auto masked_tag, const uint32_t tSb, uint64_t* bs, uint64_t* pr, const uint32_t ph, const int j,
                    const float sc, const int lim) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      ptx::mbar_wait(bs, ph, 30);
      ptx::tc_fence_after();
      uint32_t s[BN];
      tmem_ld_n<BN>(tSb, s);
      ptx::tmem_wait_ld();
      if constexpr (DBG) {
        if (p.dbg != nullptr && j * BN < 64 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
          for (int c = 0; c < BN; ++c) p.dbg[tid * 64 + j * BN + c] = (int)s[c];
        }
      }
      const int imax = row_max<BN, MASKED>(s, lim);
      const float mblk = (MASKED && imax == INT_MIN) ? -INFINITY : (float)imax * sc;
      // lazy rescale: move the reference max only when it grows by more than 2^8 (warp-uniform decision,
      // tcgen05.ld/st are warp collectives)
      if (__any_sync(0xffffffffu, mblk > m_ref + 8.f)) {
        const float m_new = fmaxf(m_ref, mblk);
        const float alpha = (m_new == -INFINITY) ? 1.f : ptx::ex2(m_ref - m_new);  // m_ref == -inf -> 0
        l *= alpha;
        m_ref = m_new;
        if (j > 0) {
          // PV_{j-1} must have landed in O.  S_j being ready implies PV_{j-2} completed (commit order) and PV_j
          // cannot start before our p_ready arrive, so bar_o is in phase j-1 or j: the parity wait is unambiguous.
          ptx::mbar_wait(bar_o, (j - 1) & 1, 31);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < D; c += 16) {
            uint32_t o[16];
            ptx::tmem_ld_x16(tOl + c, o);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st_x16(tOl + c, o);
          }
        }
      }
      uint32_t pk[BN / 2];
      l += softmax_block<BN, MASKED, VAR>(s, sc, -m_ref, lim, pk);
      tmem_st_n<BN / 2>(tSb, pk);  // P (fp16) aliases the first BN/2 columns of its S buffer
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(pr);
    };

    // blocks [0, n_full) need no mask: unrolled by two so buffer / barrier addresses are loop constants
    int n_full = nblk;
    if (CAUSAL) n_full = min(n_full, (qt * kBM) / BN);
    if (mask_tail) n_full = min(n_full, last_kblk);
    int j = 0;
    uint32_t ph = 0;
    constexpr int kPerScale = kScaleBlk / BN;  // key blocks per k_scale entry (2 for BN=32, 1 for BN=64)
    float ks_cur = ks_ptr[0];
    for (; j + 1 < n_full; j += 2, ph ^= 1) {
      const float sc0 = qs * ks_cur;
      float sc1 = sc0;
      if (kPerScale == 1) sc1 = qs * ks_ptr[j + 1];
      const float ks_nxt = ks_ptr[min((j + 2) / kPerScale, p.nkb - 1)];  // prefetch for the next pair
      step(std::false_type{}, tS0, bar_s + 0, p_ready + 0, ph, j, sc0, 0);
      step(std::false_type{}, tS1, bar_s + 1, p_ready + 1, ph, j + 1, sc1, 0);
      ks_cur = ks_nxt;
    }
    // remaining blocks (odd leftover, causal diagonal band, masked tail): generic path
    for (; j < nblk; ++j) {
      const float sc = qs * ks_ptr[min(j / kPerScale, p.nkb - 1)];
      const int c0 = j * BN;
      int lim = BN;  // columns [0, lim] are live
      if (CAUSAL && c0 + BN - 1 > qt * kBM) lim = min(lim, row - c0);
      if (mask_tail && j == last_kblk) lim = min(lim, p.Nk - 1 - c0);
      step(std::true_type{}, (j & 1) ? tS1 : tS0, bar_s + (j & 1), p_ready + (j & 1), (j >> 1) & 1, j, sc, lim);
    }

    // ---- epilogue: O / l -> out dtype, lse2 = log2(l) + m ------------------------------------------
    ptx::mbar_wait(bar_final, 0, 32);
    ptx::tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool live_row = row < p.Nq;
    uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) + ((int64_t)b * p.osb + (int64_t)hq * p.osh + (int64_t)row * p.osn) * 2;
#pragma unroll
    for (int c = 0; c < D; c += 32) {
      uint32_t o[32];
      ptx::tmem_ld_x32(tOl + c, o);  // warp collective: every lane executes it
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
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
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

template <int D, bool CAUSAL, int VAR, bool DBG = false>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                       cudaStream_t st) {
  auto kern = attn_fwd_kernel<D, CAUSAL, VAR, DBG>;
  static bool configured = false;
  if (!configured) {
    LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<D>::kBytes));
    configured = true;
  }
  dim3 grid((p.Nq + kBM - 1) / kBM, p.Hq, B);
  kern<<<grid, kThreads, AttnSmem<D>::kBytes, st>>>(tq, tk, tv, p);
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
  if (make_map(&tk, k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, B, Hkv, Nk, D, ksb, ksh, ksn, D, (D == 64 ? AttnCfg<64>::BN : AttnCfg<128>::BN), swz_qk)) return 1;
  if (make_map(&tv, v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, Hkv, Nk, D, vsb, vsh, vsn, 64, (D == 64 ? AttnCfg<64>::BN : AttnCfg<128>::BN), CU_TENSOR_MAP_SWIZZLE_128B)) return 1;

  AttnParams p;
  p.q_scale = q_scale; p.k_scale = k_scale; p.o = o; p.lse = lse;
  p.Hq = Hq; p.Hkv = Hkv; p.Nq = Nq; p.Nk = Nk;
  p.nqb = (Nq + 127) / 128; p.nkb = (Nk + 63) / 64;
  p.osb = osb; p.osh = osh; p.osn = osn;
  p.flags = flags; p.out_dtype = out_dtype; p.dbg = g_attn_debug;
  static int variant = -1;  // development switch (LOWBIT_ATTN_VARIANT): A/B of softmax instruction selection
  if (variant < 0) { const char* e = getenv("LOWBIT_ATTN_VARIANT"); variant = e ? atoi(e) : 0; }
#define LAUNCH(VAR)                                                                                              \
  if (D == 64) return causal ? launch_attn<64, true, VAR>(tq, tk, tv, p, B, st) : launch_attn<64, false, VAR>(tq, tk, tv, p, B, st); \
  return causal ? launch_attn<128, true, VAR>(tq, tk, tv, p, B, st) : launch_attn<128, false, VAR>(tq, tk, tv, p, B, st);
  if (p.dbg != nullptr && !causal)  // diagnostics build of the kernel (raw score dump)
    return D == 64 ? launch_attn<64, false, 0, true>(tq, tk, tv, p, B, st) : launch_attn<128, false, 0, true>(tq, tk, tv, p, B, st);
  if (variant == 1) { LAUNCH(1) }
  if (variant == 2) { LAUNCH(2) }
  LAUNCH(0)
#undef LAUNCH
}
