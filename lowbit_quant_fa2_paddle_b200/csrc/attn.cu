// attn.cu -- fused low-bit FlashAttention forward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   _attn_fwd / _attn_fwd_inner   src/triton/attn_qk_int8_per_block.py:24-167        (non-causal)
//   _attn_fwd_base                src/triton/attn_qk_int8_per_block_causal.py:216-334 (causal)
//   forward_merging               src/triton/quantization/attn_qk_int4_per_block.py:248-317 (INT4 K)
//   FP8 P.V semantics             csrc/qattn/qk_int_sv_f8_cuda.cu:44-692, attn_utils.cuh:30,424-428,550-562
//
// Per CTA: one 128-row Q tile of one (batch, q-head).  Q (int8), K (int8, or INT4 packed two codes per byte) and
// V (fp16, or e4m3 transposed) tiles are staged in shared memory by TMA (OOB rows zero-filled = the reference's
// masked loads); S = Q.K^T runs on tcgen05 kind::i8 (exact int32 in TMEM); each of the 128 softmax threads owns one
// row (one TMEM lane): integer row-max, dequant by q_scale*k_scale, exp2 online softmax in registers, P written back
// to TMEM as fp16 / e4m3 (aliasing S) and consumed as the A operand of the P.V tcgen05 MMA (kind::f16 or
// kind::f8f6f4) whose fp32 accumulator O stays resident in TMEM for the whole key loop (rescaled only when the
// running max moves by more than a threshold).  Epilogue: O/l (* v_scale + v_mean) -> fp16/bf16, lse2 = log2(l) + m;
// or, for ring / sequence-parallel steps, a merge into the fp32 running state (m, l, O_acc) in HBM.
//
// INT4 K: packed tiles land linearly in a staging ring; the softmax threads expand the nibbles to int8 (as
// code*16: a shift and a mask per 8 codes, no sign extension) straight into the swizzled operand layout.  The
// expansion yields head-dim order [d0 d2 d4 d6 d1 d3 d5 d7] inside every group of 8, so the Q tile is permuted the
// same way once per CTA (a contraction is invariant under a common permutation of its reduction index) and the
// factor 16 is folded into the dequantization scale (exact: a power of two).
//
// Mixed-width K (dynamic INT8 / INT4 / INT2 per 64-key block, `kbits`): K lives in a container of D bytes per row of
// which a block uses the first D*bits/8; the producer picks one of three tensor maps (box D, D/2, D/4 bytes) per
// block, so HBM / L2 traffic follows the bit width, and the softmax threads expand 8-bit (copy), 4-bit (as above) or
// 2-bit (four shift+mask per 16 codes, code*64) rows into the same permuted int8 operand layout; the quantizer
// writes 8- and 2-bit rows in the byte order that makes those expansions land in that layout (quant.cu).
#include "common.cuh"
#include "ptx.cuh"
#include "softmax_chunk.cuh"

#include <cuda.h>
#include <limits.h>
#include <stdlib.h>
#include <type_traits>

namespace lowbit {

struct AttnParams {
  const float* q_scale;
  const float* k_scale;
  const float* v_scale;  // PV e4m3: [B,Hkv,D]
  const float* v_mean;   // PV e4m3: [B,Hkv,D] or null
  const int32_t* kbits;  // mixed-width K: [B,Hkv,nkb] in {8,4,2}
  // varlen (packed [T,H,D] tensors): sequence b owns rows cu_q[b]..cu_q[b+1] (queries) / cu_k[b].. (keys); its scale
  // blocks start at cu_qs[b] / cu_ks[b] inside the head-major scale arrays [H][nqb] / [Hkv][nkb] (nqb, nkb = strides)
  const int32_t *cu_q, *cu_k, *cu_qs, *cu_ks;
  void* o;
  float* lse;
  float* m_io;           // partial (ring) state: [B,Hq,Nq]; oacc_io != null selects the merge epilogue
  float* l_io;
  float* oacc_io;        // [B,Hq,Nq,D] fp32
  int Hq, Hkv, Nq, Nk;
  int nqb, nkb;          // scale blocks per (b,h): ceil(Nq/128), ceil(Nk/64)
  int64_t osb, osh, osn;
  int delta;             // q_offset - k_offset: global position of query row 0 minus that of key 0 (causal)
  int flags, out_dtype, first;
  int csec;              // causal tile order: (batch, head) pairs per longest-first section (tile_coords)
  int32_t* dbg;          // diagnostics: when non-null, CTA (0,0,0) dumps the int32 scores of key block 0 ([128][64])
};

static int32_t* g_attn_debug = nullptr;


constexpr int kBM = 128;      // Q rows per CTA (= TMEM lanes)
constexpr int kScaleBlk = 64; // k_scale granularity of the reference quantizer (BLKK)

enum { KM_I8 = 0, KM_K4 = 1, KM_MIX = 2, KM_F16 = 3 };  // KM_K4 / KM_MIX: K tiles are expanded in shared memory;
                                                        // KM_F16: Q, K stay fp16 / bf16 (kind::f16 QK^T, fp32 scores)
enum { PV_F16 = 0, PV_E4M3 = 1 };

// Two kernels share this file:
//   attn_fwd_n64_kernel   head_dim 64, INT8 / packed-INT4 K, fp16 P.V -- 64-key steps, four CTAs per SM (the default there)
//   attn_fwd_kernel       head_dim 128 (64-key steps); head_dim 64 with the mixed-width K container or FP8 P.V (32-key steps)
//
// Per-head-dim tiling of attn_fwd_kernel.  D=64: 32-key steps keep the register footprint under 96 so that four CTAs
// (16 softmax warps) share an SM.  D=128 has twice the tensor work per exp2 and runs 64-key steps, two CTAs per SM
// (TMEM: 2 x 64 score columns + 128 output columns = 256 per CTA).
// K tiles that need expansion (packed INT4 / mixed width): a dedicated expander warp does it, so the softmax warps --
// the critical path -- carry no unpack instructions and the expansion runs ahead of them.
template <int D, int KM> struct AttnRoles {
  static constexpr bool kExpander = (KM == KM_K4 || KM == KM_MIX);
  static constexpr int kThreads = 128 + 32 + (kExpander ? 32 : 0);  // softmax warps + helper warp (+ expander warp)
};

template <int D> struct AttnCfg;
template <> struct AttnCfg<64> {
  static constexpr int BN = 32, CTAS = 4, KS = 4, VS = 3, TMEM_COLS = 128;
};
template <> struct AttnCfg<128> {
  static constexpr int BN = 64, CTAS = 2, KS = 4, VS = 3, TMEM_COLS = 256;
};

// PV-mode constants: THR = how far a block maximum may exceed the reference maximum before O is rescaled (P stays
// below 2^THR * 2^OFF: 256 in fp16, 448 in e4m3); OFF = exponent offset of the stored P (attn_utils.cuh:30 uses
// 8.807 with an exact running maximum; a lazy maximum needs THR of headroom below the e4m3 limit 448 = 2^8.807).
template <int PV> struct PvCfg;
template <> struct PvCfg<PV_F16> { static constexpr float THR = 8.f, OFF = 0.f; };
template <> struct PvCfg<PV_E4M3> { static constexpr float THR = 2.f, OFF = 6.807f; };

template <int D, int KM, int PV>
struct AttnSmem {
  using C = AttnCfg<D>;
  static constexpr int kEQ = (KM == KM_F16) ? 2 : 1;                  // bytes per Q / K operand element
  static constexpr int kQ = kBM * D * kEQ;                            // int8 (fp16 / bf16)
  static constexpr int kK = C::BN * D * kEQ;                          // operand stage
  static constexpr int kKStages = (KM == KM_I8) ? C::KS : ((KM == KM_F16 && D == 64) ? C::KS : 2);
  static constexpr int kVStages = (KM == KM_F16 && D == 128) ? 2 : C::VS;  // fp16 operands at D=128: 96 KB per CTA
  static constexpr int kKp = (KM == KM_MIX) ? C::BN * D : C::BN * D / 2;  // packed staging stage (worst case)
  static constexpr int kKpStages = (KM == KM_K4) ? 4 : (KM == KM_MIX ? 3 : 0);
  static constexpr int kV = (PV == PV_F16) ? C::BN * D * 2 : C::BN * D;  // fp16 [key][d] / e4m3 [d][key]
  static constexpr int kC = (KM == KM_F16) ? 0 : 2304;                // constant operand of the bias MMA
  static constexpr int kBytes = kQ + kKStages * kK + kVStages * kV + kKpStages * kKp + kC + 256 /*barriers*/ + 1024 /*align*/;
};

// One softmax step over a BN-key block for one query row: p = exp2(S*sc + nm), packed to fp16 pairs, row sum in fp32
// (packed fp32x2 arithmetic, FFMA2 / FADD2: one instruction scales, or accumulates, two scores).
// MASKED: columns > lim contribute 0 (causal diagonal band / masked tail keys).
// FS: 0 = the QK^T accumulator is int32, 1 = fp32 (kind::f16), 2 = int32 on top of the bias 0x4B400000 (attn_fwd_n64_kernel)
template <int FS> __device__ __forceinline__ float score_f32(uint32_t v) {
  return FS == 0 ? __int2float_rn((int)v) : __uint_as_float(v);
}
// FS == 2: (12582912 + s) - 12582912 is exact in fp32 and costs one packed add per two scores
template <int FS> __device__ __forceinline__ float2 score_pair(uint32_t a, uint32_t b) {
  const float2 v = make_float2(score_f32<FS>(a), score_f32<FS>(b));
  return FS == 2 ? __fadd2_rn(v, make_float2(-chunk::kScoreBiasF, -chunk::kScoreBiasF)) : v;
}
template <int BN, bool MASKED, int FS = 0>
__device__ __forceinline__ float softmax_block_f16(const uint32_t* __restrict__ s, float sc, float nm, int lim,
                                                   uint32_t* __restrict__ pk) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < BN; c += 4) {
    const float2 x0 = __ffma2_rn(score_pair<FS>(s[c], s[c + 1]), sc2, nm2);
    const float2 x1 = __ffma2_rn(score_pair<FS>(s[c + 2], s[c + 3]), sc2, nm2);
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
}

// FP8 variant: p~ = e4m3_rn_satfinite(exp2(S*sc + nm)) (nm carries -m + OFF), four codes per TMEM word.  The row sum
// is taken over the ROUNDED values (accumulate_d_f8, attn_utils.cuh:550-562), here through exact e4m3 -> f16
// conversion and short f16x2 partial sums.  Key c of an aligned 16-group sits at the K index the reference's V layout
// expects (fused.cu:290-292): word w of a group = keys {2w, 2w+1, 8+2w, 9+2w}.
template <int BN, bool MASKED, int FS = 0>
__device__ __forceinline__ float softmax_block_e4m3(const uint32_t* __restrict__ s, float sc, float nm, int lim,
                                                    uint32_t* __restrict__ pk) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  uint32_t hacc[BN / 16];
#pragma unroll
  for (int g = 0; g < BN / 16; ++g) {
    uint32_t h = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int c0 = 16 * g + 2 * w, c1 = c0 + 8;
      const float2 x0 = __ffma2_rn(score_pair<FS>(s[c0], s[c0 + 1]), sc2, nm2);
      const float2 x1 = __ffma2_rn(score_pair<FS>(s[c1], s[c1 + 1]), sc2, nm2);
      float2 p0 = make_float2(ptx::ex2(x0.x), ptx::ex2(x0.y));
      float2 p1 = make_float2(ptx::ex2(x1.x), ptx::ex2(x1.y));
      if (MASKED) {
        p0.x = (c0 <= lim) ? p0.x : 0.f;
        p0.y = (c0 + 1 <= lim) ? p0.y : 0.f;
        p1.x = (c1 <= lim) ? p1.x : 0.f;
        p1.y = (c1 + 1 <= lim) ? p1.y : 0.f;
      }
      const uint16_t lo = ptx::pack_e4m3x2(p0.x, p0.y), hi = ptx::pack_e4m3x2(p1.x, p1.y);
      pk[4 * g + w] = (uint32_t)lo | ((uint32_t)hi << 16);
      const uint32_t hs = ptx::hadd2(ptx::e4m3x2_to_f16x2(lo), ptx::e4m3x2_to_f16x2(hi));
      h = (w == 0) ? hs : ptx::hadd2(h, hs);
    }
    hacc[g] = h;  // 4 values per f16 lane, each <= 448: no overflow, relative error <= 2^-10
  }
  float t = 0.f;
#pragma unroll
  for (int g = 0; g < BN / 16; ++g) t += ptx::f16x2_sum(hacc[g]);
  return t;
}

template <int N> __device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 16) ptx::tmem_ld_x16(taddr, r);
  else if constexpr (N == 32) ptx::tmem_ld_x32(taddr, r);
  else { ptx::tmem_ld_x32(taddr, r); ptx::tmem_ld_x32(taddr + 32, r + 32); }
}
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
  if constexpr (N == 4) ptx::tmem_st_x4(taddr, r);
  else if constexpr (N == 8) ptx::tmem_st_x8(taddr, r);
  else if constexpr (N == 16) ptx::tmem_st_x16(taddr, r);
  else ptx::tmem_st_x32(taddr, r);
}

// INT4 K: expand one packed staging tile (BN rows x D/2 bytes, linear) into the int8 operand stage (BN rows x D
// bytes, K-major, 64B / 128B hardware swizzle).  Every 16-byte packed chunk (32 codes) becomes two 16-byte chunks.
template <int D, int BN, int NT>
__device__ __forceinline__ void unpack_k4_tile(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int tid) {
  constexpr int kChunks = BN * D / 32;  // packed 16-byte chunks in the tile
  constexpr int kCPR = D / 32;          // packed chunks per row
  constexpr uint32_t kSwzMask = (D == 64) ? 3u : 7u;
  constexpr uint32_t kM = 0xF0F0F0F0u;
#pragma unroll
  for (int q = tid; q < kChunks; q += NT) {
    const uint4 w = *reinterpret_cast<const uint4*>(src + q * 16);
    const uint32_t off0 = (uint32_t)(q / kCPR) * D + (uint32_t)(q % kCPR) * 32, off1 = off0 + 16;
    uint4 a, b;
    a.x = (w.x << 4) & kM; a.y = w.x & kM; a.z = (w.y << 4) & kM; a.w = w.y & kM;
    b.x = (w.z << 4) & kM; b.y = w.z & kM; b.z = (w.w << 4) & kM; b.w = w.w & kM;
    *reinterpret_cast<uint4*>(dst + (off0 ^ (((off0 >> 7) & kSwzMask) << 4))) = a;
    *reinterpret_cast<uint4*>(dst + (off1 ^ (((off1 >> 7) & kSwzMask) << 4))) = b;
  }
}
// mixed-width K, 8-bit block: rows are already in the permuted byte order; move 16-byte chunks into the swizzled layout
template <int D, int BN, int NT>
__device__ __forceinline__ void copy_k8_tile(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int tid) {
  constexpr uint32_t kSwzMask = (D == 64) ? 3u : 7u;
#pragma unroll
  for (int q = tid; q < BN * D / 16; q += NT) {
    const uint32_t off = (uint32_t)q * 16;
    *reinterpret_cast<uint4*>(dst + (off ^ (((off >> 7) & kSwzMask) << 4))) = *reinterpret_cast<const uint4*>(src + off);
  }
}
// mixed-width K, 2-bit block: one packed 16-byte chunk (64 codes, field k of byte i = bits [2k, 2k+2)) becomes four
// 16-byte chunks of code*64 (field k of every byte -> output chunk k)
template <int D, int BN, int NT>
__device__ __forceinline__ void unpack_k2_tile(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int tid) {
  constexpr int kChunks = BN * D / 64;
  constexpr int kCPR = D / 64;
  constexpr uint32_t kSwzMask = (D == 64) ? 3u : 7u;
  constexpr uint32_t kM = 0xC0C0C0C0u;
#pragma unroll
  for (int q = tid; q < kChunks; q += NT) {
    const uint4 w = *reinterpret_cast<const uint4*>(src + q * 16);
    const uint32_t off0 = (uint32_t)(q / kCPR) * D + (uint32_t)(q % kCPR) * 64;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 o;
      o.x = (w.x << (6 - 2 * k)) & kM; o.y = (w.y << (6 - 2 * k)) & kM;
      o.z = (w.z << (6 - 2 * k)) & kM; o.w = (w.w << (6 - 2 * k)) & kM;
      const uint32_t off = off0 + 16 * k;
      *reinterpret_cast<uint4*>(dst + (off ^ (((off >> 7) & kSwzMask) << 4))) = o;
    }
  }
}
// expand one staged K tile of the given bit width into the int8 operand stage
template <int D, int BN, int NT, int KM>
__device__ __forceinline__ void expand_k_tile(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int tid, int bits) {
  if constexpr (KM == KM_MIX) {
    if (bits == 8) copy_k8_tile<D, BN, NT>(src, dst, tid);
    else if (bits == 4) unpack_k4_tile<D, BN, NT>(src, dst, tid);
    else unpack_k2_tile<D, BN, NT>(src, dst, tid);
  } else {
    unpack_k4_tile<D, BN, NT>(src, dst, tid);
  }
}
// the matching permutation of the Q tile: [d0..d7] -> [d0 d2 d4 d6 d1 d3 d5 d7] inside every 8-byte group
// (16-byte chunks move as units under the hardware swizzle, so the tile can be walked linearly)
template <int D, int NT>
__device__ __forceinline__ void permute_q_tile(uint8_t* sQ, int tid) {
#pragma unroll
  for (int q = tid; q < kBM * D / 16; q += NT) {
    uint4 w = *reinterpret_cast<uint4*>(sQ + q * 16), r;
    r.x = __byte_perm(w.x, w.y, 0x6420); r.y = __byte_perm(w.x, w.y, 0x7531);
    r.z = __byte_perm(w.z, w.w, 0x6420); r.w = __byte_perm(w.z, w.w, 0x7531);
    *reinterpret_cast<uint4*>(sQ + q * 16) = r;
  }
}


// Row epilogue of both kernels (thread <-> query row <-> TMEM lane; tOl = this warp's lanes of the fp32 accumulator):
//   O / l (* v_scale + v_mean) -> fp16 / bf16, lse2 = log2(l) + m - OFF;
//   or, ring steps (oacc_io != null): merge (m_ref, l, O) of this K/V shard into the running fp32 state in HBM.
// vsc / vmn: per-channel V scale / mean of this (b, kv head), staged in SHARED memory at kernel start -- read straight
// from global memory here, every element's load sat behind the previous store to the (possibly aliasing) fp32 state
// and cost the ring merge epilogue 0.76 ms per 8K x 8K launch.
template <int D, int PV>
__device__ __forceinline__ void attn_epilogue(const AttnParams& p, const uint32_t tOl, const float l, const float m_ref,
                                              const int row, const int Nq, const int b, const int hq,
                                              const int64_t orow_base, const float* vsc, const float* vmn) {
  using PC = PvCfg<PV>;
  const bool live_row = row < Nq;
  const int64_t idx = ((int64_t)b * p.Hq + hq) * Nq + row;
  if (p.oacc_io == nullptr) {
    const float inv_l = 1.0f / l;
    uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) + (orow_base + (int64_t)hq * p.osh + (int64_t)row * p.osn) * 2;
#pragma unroll
    for (int c = 0; c < D; c += 32) {
      uint32_t o[32];
      ptx::tmem_ld_x32(tOl + c, o);  // warp collective: every lane executes it
      ptx::tmem_wait_ld();
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(o[2 * i]) * inv_l, bb = __uint_as_float(o[2 * i + 1]) * inv_l;
        if constexpr (PV == PV_E4M3) {
          a *= vsc[c + 2 * i];
          bb *= vsc[c + 2 * i + 1];
          if (vmn) { a += vmn[c + 2 * i]; bb += vmn[c + 2 * i + 1]; }
        }
        w[i] = (p.out_dtype == LOWBIT_F16) ? ptx::pack_f16x2(a, bb) : ptx::pack_bf16x2(a, bb);
      }
      if (live_row) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
    }
    if (live_row && p.lse) p.lse[idx] = ptx::lg2(l) + m_ref - PC::OFF;
  } else {
    // in true (dequantized) units
    float m_prev = -INFINITY, l_prev = 0.f;
    if (!p.first && live_row) { m_prev = p.m_io[idx]; l_prev = p.l_io[idx]; }
    const float m_cur = (l > 0.f) ? m_ref : -INFINITY;  // a row that saw only masked keys contributes nothing
    const float m_new = fmaxf(m_prev, m_cur);
    const float wa = (m_prev == -INFINITY) ? 0.f : ptx::ex2(m_prev - m_new);
    const float wb0 = (m_cur == -INFINITY) ? 0.f : ptx::ex2(m_cur - m_new);
    const float wb = wb0 * ((PV == PV_E4M3) ? exp2f(-PC::OFF) : 1.f);  // stored P carries 2^OFF
    const float l_cur = l * wb;
    float* od = p.oacc_io + idx * D;
#pragma unroll
    for (int c = 0; c < D; c += 16) {
      uint32_t o[16];
      ptx::tmem_ld_x16(tOl + c, o);
      ptx::tmem_wait_ld();
      if (live_row) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!p.first) prev = *reinterpret_cast<const float4*>(od + c + i);
          float v[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float cur = __uint_as_float(o[i + t]) * wb;
            if constexpr (PV == PV_E4M3) {
              cur *= vsc[c + i + t];
              if (vmn) cur += vmn[c + i + t] * l_cur;
            }
            v[t] = cur;
          }
          prev.x = prev.x * wa + v[0]; prev.y = prev.y * wa + v[1];
          prev.z = prev.z * wa + v[2]; prev.w = prev.w * wa + v[3];
          *reinterpret_cast<float4*>(od + c + i) = prev;
        }
      }
    }
    if (live_row) {
      p.m_io[idx] = m_new;
      p.l_io[idx] = l_prev * wa + l_cur;
    }
  }
}

// What one CTA works on: the padded tensors' (b, h) slice, or -- varlen -- sequence b of the packed tensors.
struct TileView {
  int Nq, Nk, nkb;          // rows of this (b, h) / sequence; k_scale blocks
  int q_row0, k_row0, tb;   // row offsets into the packed tensors; batch coordinate of the tensor maps
  int64_t qs_idx, ks_base;  // my q_scale entry; first k_scale / kbits entry of my (b, kv head)
  int64_t orow_base;        // element offset of my first output row (head added later)
  bool done;                // nothing left to do for this CTA
};

// Which (Q tile, head, batch) a CTA works on.  Non-causal: the grid coordinates.  Causal: a tile's work grows with its
// index (2 .. 2 nqt key blocks), and CTAs are dispatched in linear block order, so the order is longest-first over
// sections of p.csec (batch, head) pairs (32; LOWBIT_CAUSAL_SECTION) -- all their heaviest tiles, then the next lighter ones ... -- instead of
// longest-first inside each (batch, head): the grid then ends on the lightest tiles of the last section (a few key
// blocks each) rather than on heavy tiles of the last heads that started late, and a section's K / V (kSec x
// (N x D x 3 bytes)) stays L2-resident while its tiles run.
__device__ __forceinline__ void tile_coords(const bool causal, const int kCausalSection, int& qt, int& hq, int& b) {
  if (!causal) {
    qt = blockIdx.x; hq = blockIdx.y; b = blockIdx.z;
    return;
  }
  const int nqt = gridDim.x, nbh = gridDim.y * gridDim.z;
  const int id = blockIdx.x + nqt * (blockIdx.y + gridDim.y * blockIdx.z);
  const int sec = id / (kCausalSection * nqt), base = sec * kCausalSection;
  const int cnt = min(kCausalSection, nbh - base), r = id - base * nqt;
  qt = nqt - 1 - r / cnt;
  const int bh = base + r % cnt;
  hq = bh % (int)gridDim.y;
  b = bh / (int)gridDim.y;
}

// Resolves the CTA's view and handles the degenerate tiles (past the end of a varlen sequence; a sequence without keys).
template <int D>
__device__ __forceinline__ TileView tile_view(const AttnParams& p, const int qt, const int hq, const int hkv, const int b,
                                              const int tid) {
  TileView t;
  t.Nq = p.Nq; t.Nk = p.Nk; t.nkb = p.nkb;
  t.q_row0 = 0; t.k_row0 = 0; t.tb = b;
  t.qs_idx = ((int64_t)b * p.Hq + hq) * p.nqb + qt;
  t.ks_base = ((int64_t)b * p.Hkv + hkv) * p.nkb;
  t.orow_base = (int64_t)b * p.osb;
  t.done = false;
  if (p.cu_q != nullptr) {
    t.q_row0 = p.cu_q[b];
    t.Nq = p.cu_q[b + 1] - t.q_row0;
    t.k_row0 = p.cu_k[b];
    t.Nk = p.cu_k[b + 1] - t.k_row0;
    if (qt * kBM >= t.Nq) { t.done = true; return t; }  // the grid is sized for the longest sequence (attn_qk_int8_block_varlen.py:126-128)
    t.nkb = (t.Nk + kScaleBlk - 1) / kScaleBlk;
    t.qs_idx = (int64_t)hq * p.nqb + p.cu_qs[b] + qt;
    t.ks_base = (int64_t)hkv * p.nkb + p.cu_ks[b];
    t.orow_base = (int64_t)t.q_row0 * p.osn;
    t.tb = 0;
    if (t.Nk == 0) {  // no keys: the reference divides a zero accumulator by l = 1 (attn_qk_int8_block_varlen.py:168-189)
      if (tid < kBM && qt * kBM + tid < t.Nq) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.o) +
                                              (t.orow_base + (int64_t)hq * p.osh + (int64_t)(qt * kBM + tid) * p.osn) * 2);
#pragma unroll
        for (int i = 0; i < D / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
      }
      t.done = true;
    }
  }
  return t;
}

// ring step whose K/V shard lies wholly in this tile's future: the running state is unchanged (initialised when first)
template <int D>
__device__ __forceinline__ void empty_ring_step(const AttnParams& p, const int qt, const int hq, const int b,
                                                const int Nq, const int tid) {
  if (p.oacc_io != nullptr && p.first && tid < kBM) {
    const int row = qt * kBM + tid;
    if (row < Nq) {
      const int64_t idx = ((int64_t)b * p.Hq + hq) * Nq + row;
      p.m_io[idx] = -INFINITY;
      p.l_io[idx] = 0.f;
      float4* od = reinterpret_cast<float4*>(p.oacc_io + idx * D);
#pragma unroll
      for (int i = 0; i < D / 4; ++i) od[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Warp roles:  warps 0-3  softmax (thread t <-> query row t <-> TMEM lane t)
//              warp 4     helper: one elected lane is both the TMA producer and the tcgen05 issuer; the warp also
//                         owns the TMEM allocation
//              warp 5     (packed / mixed-width K) expander: packed K tiles -> int8 operand stages
// TMEM columns: S/P buffer 0 [0,BN)  S/P buffer 1 [BN,2BN)  O [2BN, 2BN+D)
// Pipeline: QK_{j+2} is issued right after PV_j, so the int8 contraction of the next two key blocks and the
// P.V of the previous one run on the tensor pipe while the softmax warps work on block j; K/V stages are refilled
// by the same thread as soon as the MMAs that read them have committed.
template <int D, int KM, int PV, bool DBG>
__global__ void __launch_bounds__((AttnRoles<D, KM>::kThreads), AttnCfg<D>::CTAS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK8,
                const __grid_constant__ CUtensorMap tmK2, const AttnParams p) {
  using C = AttnCfg<D>;
  using SM = AttnSmem<D, KM, PV>;
  using PC = PvCfg<PV>;
  constexpr int BN = C::BN, VS = SM::kVStages;
  constexpr int KS = SM::kKStages;                   // int8 operand stages (expanded K: 2, indexed like the S buffers)
  constexpr bool KX = (KM == KM_K4 || KM == KM_MIX);          // K tiles are expanded by the expander warp
  constexpr bool FQK = (KM == KM_F16);                        // fp16 / bf16 operands, fp32 scores
  // int8 scores are accumulated on top of 0x4B400000 written by a small fp16 MMA: they leave TMEM as the fp32 number
  // 12582912 + s and the softmax needs no int -> float conversion (attn_fwd_n64_kernel, softmax_chunk.cuh); exact
  // while |s| < 2^22 (head_dim 128: |s| <= 2^21)
  constexpr bool BIASED = !FQK;
  constexpr int KPS = KX ? SM::kKpStages : KS;                // TMA-filled K stages
  constexpr int PCOLS = (PV == PV_F16) ? BN / 2 : BN / 4;     // TMEM columns of one P tile
  constexpr int HW = 4;                                       // index of the helper warp
  // p_ready counts warps, not threads (measured: D=128 causal +3 %, D=64 INT8 K unchanged, but -1.8 % on the D=64
  // expander-warp kernels, which keep per-thread arrivals)
  constexpr bool kWarpArrive = !(D == 64 && KX);
  __shared__ float s_vs[PV == PV_E4M3 ? D : 1], s_vm[PV == PV_E4M3 ? D : 1];  // FP8 P.V: v_scale / v_mean of this (b, kv head)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SM::kQ;
  uint8_t* sV = sK + KS * SM::kK;
  uint8_t* sKp = sV + VS * SM::kV;
  uint8_t* sC = sKp + SM::kKpStages * SM::kKp;  // bias-MMA operand (see attn_fwd_n64_kernel)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + SM::kC);
  uint64_t* bar_q = bars + 0;
  uint64_t* kfull = bars + 1;        // [KPS] TMA -> consumer (MMA issuer, or the expander warp)
  uint64_t* kfree = kfull + KPS;     // [KPS] consumer -> TMA
  uint64_t* vfull = kfree + KPS;     // [VS]
  uint64_t* vfree = vfull + VS;      // [VS]
  uint64_t* bar_s = vfree + VS;      // [2] QK done: S buffer b holds scores
  uint64_t* p_ready = bar_s + 2;     // [2] the softmax warps wrote P into buffer b
  uint64_t* bar_o = p_ready + 2;     // PV_j done (one phase per key block)
  uint64_t* bar_final = bar_o + 1;   // last PV done (single phase: parity waits must never lag 2 phases)
  uint64_t* bar_k01 = bar_final + 1; // expander warp: Q permuted
  uint64_t* kready = bar_k01 + 1;    // [2] expander warp: operand stage holds the expanded K tile
  uint64_t* kopfree = kready + 2;    // [2] QK on an operand stage complete: the expander may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kopfree + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool causal = (p.flags & LOWBIT_ATTN_CAUSAL) != 0;
  int qt, hq, b;
  tile_coords(causal, p.csec, qt, hq, b);
  const int hkv = hq / (p.Hq / p.Hkv);

  const TileView tv = tile_view<D>(p, qt, hq, hkv, b, tid);
  if (tv.done) return;
  const int Nq = tv.Nq, Nk = tv.Nk, nkb = tv.nkb, q_row0 = tv.q_row0, k_row0 = tv.k_row0, tb = tv.tb;
  const int64_t qs_idx = tv.qs_idx, ks_base = tv.ks_base, orow_base = tv.orow_base;

  // key-block range of this Q tile.  compat_tail walks the reference's whole 64-key blocks (phantom zero keys).
  const bool compat = (p.flags & LOWBIT_ATTN_COMPAT_TAIL) != 0;
  const int nk_eff = compat ? nkb * kScaleBlk : Nk;
  int nblk = (nk_eff + BN - 1) / BN;
  // causal: key (global k_off + c) is visible to row (global q_off + r) iff c <= delta + r
  const int dq = p.delta + qt * kBM;  // delta + first row of the tile
  if (causal) nblk = max(0, min(nblk, (dq + kBM + BN - 1) / BN));
  if (nblk == 0) {
    empty_ring_step<D>(p, qt, hq, b, Nq, tid);
    return;
  }

  if (warp == HW) {
    ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (tid == 0) {
    ptx::mbar_init(bar_q, 1);
    for (int i = 0; i < KPS; ++i) {
      ptx::mbar_init(kfull + i, 1);
      ptx::mbar_init(kfree + i, KX ? 32 : 1);
    }
    for (int i = 0; i < VS; ++i) { ptx::mbar_init(vfull + i, 1); ptx::mbar_init(vfree + i, 1); }
    // p_ready: one arrival per softmax warp (tcgen05.wait::st is a warp collective, so lane 0 can speak for the warp)
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_s + i, 1); ptx::mbar_init(p_ready + i, kWarpArrive ? 4 : 128); }
    ptx::mbar_init(bar_o, 1);
    ptx::mbar_init(bar_final, 1);
    ptx::mbar_init(bar_k01, 32);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(kready + i, 32); ptx::mbar_init(kopfree + i, 1); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    if constexpr (KM == KM_MIX) { ptx::prefetch_tmap(&tmK8); ptx::prefetch_tmap(&tmK2); }
  }
  if constexpr (PV == PV_E4M3) {
    if (tid < D) {
      s_vs[tid] = p.v_scale[((int64_t)b * p.Hkv + hkv) * D + tid];
      s_vm[tid] = p.v_mean ? p.v_mean[((int64_t)b * p.Hkv + hkv) * D + tid] : 0.f;
    }
  }
  if constexpr (BIASED) {  // every 16-byte row chunk = fp16 {1024 x 6, 0 x 2}: 12 * 2^20 = 12582912.0f = 0x4B400000 per score
    if (tid < SM::kC / 16) ptx::sts_v4_a(ptx::smem_u32(sC) + tid * 16, 0x64006400u, 0x64006400u, 0x64006400u, 0u);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 2 * BN;  // fp32 output accumulator, D columns
  const int32_t* kb_ptr = (KM == KM_MIX) ? p.kbits + ks_base : nullptr;
  auto kb = [&](int jj) -> int {  // bit width of key block jj
    if constexpr (KM == KM_MIX) return kb_ptr[min(jj * BN / kScaleBlk, nkb - 1)];
    else return 4;
  };

  if (warp == HW) {
    // ================================ helper: TMA producer + tcgen05 issuer ================================
    if (ptx::elect_one()) {
      constexpr uint32_t kSwzQK = (D == 64) ? ptx::kSwz64 : ptx::kSwz128;
      constexpr uint32_t kSboQK = 8 * D;  // 8 rows of D bytes
      const uint32_t qk_fmt = (p.out_dtype == LOWBIT_F16) ? ptx::kF16 : ptx::kBF16;  // KM_F16: q, k in the output dtype
      const uint32_t idesc_qk = FQK ? ptx::make_idesc(ptx::kCF32, qk_fmt, qk_fmt, 0, 0, kBM, BN)
                                    : ptx::make_idesc(ptx::kCS32, ptx::kS8, ptx::kS8, 0, 0, kBM, BN);
      constexpr uint32_t idesc_pv = (PV == PV_F16) ? ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 1, kBM, D)
                                                   : ptx::make_idesc(ptx::kCF32, ptx::kE4M3, ptx::kE4M3, 0, 0, kBM, D);
      const uint32_t aq = ptx::smem_u32(sQ);
      auto load_k = [&](int j) {  // int8 tile (swizzled) or packed tile (linear) into TMA stage j % KPS
        const int ks = j % KPS;
        ptx::mbar_wait(kfree + ks, ((j / KPS) & 1) ^ 1, 10);
        if constexpr (FQK) {  // fp16 rows: 64-element (128-byte) swizzle atoms of BN rows each
          ptx::mbar_expect_tx(kfull + ks, SM::kK);
#pragma unroll
          for (int a = 0; a < D / 64; ++a)
            ptx::tma_load_4d(sK + ks * SM::kK + a * BN * 128, &tmK, kfull + ks, 64 * a, k_row0 + j * BN, hkv, tb);
        } else if constexpr (KM == KM_I8) {
          ptx::mbar_expect_tx(kfull + ks, SM::kK);
          ptx::tma_load_4d(sK + ks * SM::kK, &tmK, kfull + ks, 0, k_row0 + j * BN, hkv, tb);
        } else if constexpr (KM == KM_K4) {
          ptx::mbar_expect_tx(kfull + ks, SM::kKp);
          ptx::tma_load_4d(sKp + ks * SM::kKp, &tmK, kfull + ks, 0, k_row0 + j * BN, hkv, tb);
        } else {  // mixed width: the block's bit width picks the box (first D*bits/8 bytes of every container row)
          const int bits = p.kbits[ks_base + min(j * BN / kScaleBlk, nkb - 1)];
          const CUtensorMap* tm = (bits == 8) ? &tmK8 : (bits == 4 ? &tmK : &tmK2);
          ptx::mbar_expect_tx(kfull + ks, BN * D * bits / 8);
          ptx::tma_load_4d(sKp + ks * SM::kKp, tm, kfull + ks, 0, k_row0 + j * BN, hkv, tb);
        }
      };
      auto load_v = [&](int j) {
        const int vs = j % VS;
        ptx::mbar_wait(vfree + vs, ((j / VS) & 1) ^ 1, 11);
        ptx::mbar_expect_tx(vfull + vs, SM::kV);
        if constexpr (PV == PV_F16) {
          ptx::tma_load_4d(sV + vs * SM::kV, &tmV, vfull + vs, 0, k_row0 + j * BN, hkv, tb);
          if (D == 128) ptx::tma_load_4d(sV + vs * SM::kV + BN * 128, &tmV, vfull + vs, 64, k_row0 + j * BN, hkv, tb);
        } else {
          ptx::tma_load_4d(sV + vs * SM::kV, &tmV, vfull + vs, j * BN, 0, hkv, b);  // [d][key] tile, keys contiguous
        }
      };
      auto issue_qk = [&](int j) {
        const int ks = j % KS;
        if constexpr (!KX) {
          ptx::mbar_wait(kfull + ks, (j / KS) & 1, 20);
        } else {
          ptx::mbar_wait(kready + ks, (j / KS) & 1, 24);  // expanded by the expander warp; its packed stage is free
          if (j + KPS < nblk) load_k(j + KPS);
        }
        ptx::tc_fence_after();
        const uint32_t ak = ptx::smem_u32(sK + ks * SM::kK);
        const uint32_t tS = tmem_base + (j & 1) * BN;
        if constexpr (FQK) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {  // K = 16 elements = 32 bytes per MMA; 4 per 128-byte swizzle atom
            const uint32_t off = (kk % 4) * 32;
            const uint64_t da = ptx::make_smem_desc(aq + (kk / 4) * kBM * 128 + off, 16, 1024, ptx::kSwz128);
            const uint64_t db = ptx::make_smem_desc(ak + (kk / 4) * BN * 128 + off, 16, 1024, ptx::kSwz128);
            ptx::umma_f16_ss(tS, da, db, idesc_qk, kk > 0);
          }
        } else {
          constexpr uint32_t idesc_bias = ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 0, kBM, BN);
          const uint64_t dc = ptx::make_smem_desc(ptx::smem_u32(sC), 128, 128, ptx::kSwzNone);
          ptx::umma_f16_ss(tS, dc, dc, idesc_bias, 0);  // every score column = 0x4B400000
#pragma unroll
          for (int kk = 0; kk < D / 32; ++kk) {
            const uint64_t da = ptx::make_smem_desc(aq + kk * 32, 16, kSboQK, kSwzQK);
            const uint64_t db = ptx::make_smem_desc(ak + kk * 32, 16, kSboQK, kSwzQK);
            ptx::umma_i8_ss(tS, da, db, idesc_qk, 1);
          }
        }
        ptx::umma_commit(bar_s + (j & 1));  // scores ready for the softmax warps
        if constexpr (!KX) ptx::umma_commit(kfree + ks);  // K stage may be refilled
        else ptx::umma_commit(kopfree + ks);               // operand stage may be overwritten
      };
      ptx::mbar_expect_tx(bar_q, SM::kQ);
      if constexpr (FQK) {
#pragma unroll
        for (int a = 0; a < D / 64; ++a) ptx::tma_load_4d(sQ + a * kBM * 128, &tmQ, bar_q, 64 * a, q_row0 + qt * kBM, hq, tb);
      } else {
        ptx::tma_load_4d(sQ, &tmQ, bar_q, 0, q_row0 + qt * kBM, hq, tb);
      }
      for (int j = 0; j < min(KPS, nblk); ++j) load_k(j);
      for (int j = 0; j < min(VS - 1, nblk); ++j) load_v(j);
      if constexpr (!KX) ptx::mbar_wait(bar_q, 0, 21);
      else ptx::mbar_wait(bar_k01, 0, 21);  // Q permuted
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        const int vs = j % VS;
        ptx::mbar_wait(p_ready + (j & 1), (j >> 1) & 1, 22);
        ptx::mbar_wait(vfull + vs, (j / VS) & 1, 23);
        ptx::tc_fence_after();
        const uint32_t av = ptx::smem_u32(sV + vs * SM::kV);
        const uint32_t tP = tmem_base + (j & 1) * BN;
        if constexpr (PV == PV_F16) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            // V tile: MN-major (d contiguous), 128B swizzle: 8 key rows = 1024 B (SBO); 64-wide d atoms BN*128 B apart (LBO)
            const uint64_t db = ptx::make_smem_desc(av + kk * 16 * 128, BN * 128, 1024, ptx::kSwz128);
            ptx::umma_f16_ts(tO, tP + kk * 8, db, idesc_pv, (j > 0) || (kk > 0));
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < BN / 32; ++kk) {
            // V^T tile: K-major (keys contiguous), rows of BN bytes (32B / 64B swizzle), 8 channel rows = 8*BN B (SBO)
            const uint64_t db = ptx::make_smem_desc(av + kk * 32, 16, 8 * BN, BN == 32 ? ptx::kSwz32 : ptx::kSwz64);
            ptx::umma_f8_ts(tO, tP + kk * 8, db, idesc_pv, (j > 0) || (kk > 0));
          }
        }
        ptx::umma_commit(vfree + vs);
        ptx::umma_commit(bar_o);
        if (j == nblk - 1) ptx::umma_commit(bar_final);
        // two K stages only (fp16 operands at D=128): K_{j+2} goes into the stage QK_j read (complete: S_j was
        // consumed) BEFORE QK_{j+2} waits for it
        if constexpr (!KX && KPS == 2) {
          if (j + 2 < nblk) load_k(j + 2);
        }
        if (j + 2 < nblk) issue_qk(j + 2);  // overwrites S/P buffer (j&1): ordered after PV_j on the tensor pipe
        // refill: the K stage consumed longest ago and the V stage of PV_{j-1} (complete in steady state)
        if constexpr (!KX && KPS > 2) {
          if (j + KPS < nblk) load_k(j + KPS);
        }
        if (j + VS - 1 < nblk) load_v(j + VS - 1);
      }
    }
  } else if (KX && warp == HW + 1) {
    // ================================ expander warp: packed K tiles -> int8 operand stages ================================
    const int et = tid & 31;
    ptx::mbar_wait(bar_q, 0, 33);
    permute_q_tile<D, 32>(sQ, et);
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(bar_k01);
    for (int j = 0; j < nblk; ++j) {
      const int ks = j % KS, kps = j % KPS;
      if (j >= KS) ptx::mbar_wait(kopfree + ks, ((j / KS) - 1) & 1, 36);  // QK_{j-KS} has read this operand stage
      ptx::mbar_wait(kfull + kps, (j / KPS) & 1, 34);
      expand_k_tile<D, BN, 32, KM>(sKp + kps * SM::kKp, sK + ks * SM::kK, et, kb(j));
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(kfree + kps);
      ptx::mbar_arrive(kready + ks);
    }
  } else {
    // ================================ softmax warps ================================
    const int r = tid & (kBM - 1);                // query row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;  // TMEM lane quadrant of this warp
    const int row = qt * kBM + r;  // query row owned by this thread
    // KM_F16: one factor for every score, sm_scale * log2(e), handed over as q_scale[0]; k_scale is not read
    float qs = FQK ? p.q_scale[0] : p.q_scale[qs_idx];
    if (KM == KM_K4) qs *= 0.0625f;  // K operand holds code*16
    const float* ks_ptr = FQK ? p.q_scale : p.k_scale + ks_base;
    auto kfac = [&](int jj) -> float {  // the expanded operand holds code, code*16 or code*64
      if constexpr (KM == KM_MIX) { const int bb = kb(jj); return bb == 8 ? 1.f : (bb == 4 ? 0.0625f : 0.015625f); }
      else return 1.f;
    };
    const bool mask_tail = !compat && (Nk % BN != 0);
    const int last_kblk = (Nk + BN - 1) / BN - 1;
    const uint32_t tS0 = tmem_base + lane_off, tS1 = tS0 + BN;  // my score columns in S buffer 0 / 1 (P aliases S)
    const uint32_t tOl = tO + lane_off;
    float m_ref = -INFINITY, l = 0.f;
    constexpr float kScMax = 1.0e-3f;

    // one key block: wait for S, row max, (rare) rescale of O, P = exp2(S*sc - m) -> TMEM, signal the issuer
    auto step = [&](auto masked_tag, const uint32_t tSb, uint64_t* bs, uint64_t* pr, const uint32_t ph, const int j,
                    const float sc_in, const int lim) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const float sc = sc_in * kfac(j);
      ptx::mbar_wait(bs, ph, 30);
      ptx::tc_fence_after();
      uint32_t s[BN];
      tmem_ld_n<BN>(tSb, s);
      ptx::tmem_wait_ld();
      if constexpr (DBG) {
        if (p.dbg != nullptr && j * BN < 64 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
          for (int c = 0; c < BN; ++c) p.dbg[r * 64 + j * BN + c] = (int)(s[c] - (BIASED ? chunk::kScoreBias : 0u));
        }
      }
      float mblk;
      if constexpr (FQK) {
        float fm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < BN; ++c) fm[c & 3] = fmaxf(fm[c & 3], (!MASKED || c <= lim) ? __uint_as_float(s[c]) : -INFINITY);
        mblk = fmaxf(fmaxf(fm[0], fm[1]), fmaxf(fm[2], fm[3])) * sc;  // sc > 0; -inf stays -inf
      } else {
        const int imax = chunk::row_max_i<BN, MASKED>(s, lim);  // biased scores keep their order as int32
        mblk = (MASKED && imax == INT_MIN) ? -INFINITY : (float)(imax - (int)(BIASED ? chunk::kScoreBias : 0u)) * sc;
      }
      // lazy rescale: move the reference max only when it grows by more than 2^THR (warp-uniform decision,
      // tcgen05.ld/st are warp collectives)
      if (__any_sync(0xffffffffu, mblk > m_ref + PC::THR)) {
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
      uint32_t pk[PCOLS];
      const float nm = (m_ref == -INFINITY) ? 0.f : PC::OFF - m_ref;  // fully masked row so far: p is zeroed by the mask
      if constexpr (FQK) {
        if constexpr (PV == PV_F16) l += softmax_block_f16<BN, MASKED, 1>(s, sc, nm, lim, pk);
        else l += softmax_block_e4m3<BN, MASKED, 1>(s, sc, nm, lim, pk);
      } else if (sc > kScMax) {  // coarse scale (warp-uniform): subtract the bias exactly (one packed add per pair)
        if constexpr (PV == PV_F16) l += softmax_block_f16<BN, MASKED, 2>(s, sc, nm, lim, pk);
        else l += softmax_block_e4m3<BN, MASKED, 2>(s, sc, nm, lim, pk);
      } else {  // the bias leaves through the FMA addend: common factor <= 2^(2^-25 * 12582912 * sc) on the block's P
        const float nmb = fmaf(-chunk::kScoreBiasF, sc, nm);
        if constexpr (PV == PV_F16) l += softmax_block_f16<BN, MASKED, 1>(s, sc, nmb, lim, pk);
        else l += softmax_block_e4m3<BN, MASKED, 1>(s, sc, nmb, lim, pk);
      }
      tmem_st_n<PCOLS>(tSb, pk);  // P aliases the first columns of its S buffer
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      if constexpr (kWarpArrive) {
        if ((tid & 31) == 0) ptx::mbar_arrive(pr);
      } else {
        ptx::mbar_arrive(pr);
      }
    };

    constexpr int kPerScale = kScaleBlk / BN;  // key blocks per k_scale entry (2 for BN=32, 1 for BN=64)
    int n_full = nblk;  // blocks [0, n_full) need no mask
    if (causal) n_full = max(0, min(n_full, (dq + 1) / BN));
    if (mask_tail) n_full = min(n_full, last_kblk);
    {
    // unrolled by two so buffer / barrier addresses are loop constants
    int j = 0;
    uint32_t ph = 0;
    float ks_cur = FQK ? 1.f : ks_ptr[0];
    for (; j + 1 < n_full; j += 2, ph ^= 1) {
      const float sc0 = qs * ks_cur;
      float sc1 = sc0;
      if (kPerScale == 1 && !FQK) sc1 = qs * ks_ptr[j + 1];
      const float ks_nxt = FQK ? 1.f : ks_ptr[min((j + 2) / kPerScale, nkb - 1)];  // prefetch for the next pair
      step(std::false_type{}, tS0, bar_s + 0, p_ready + 0, ph, j, sc0, 0);
      step(std::false_type{}, tS1, bar_s + 1, p_ready + 1, ph, j + 1, sc1, 0);
      ks_cur = ks_nxt;
    }
    // remaining blocks (odd leftover, causal diagonal band, masked tail): generic path
    for (; j < nblk; ++j) {
      const float sc = FQK ? qs : qs * ks_ptr[min(j / kPerScale, nkb - 1)];
      const int c0 = j * BN;
      int lim = BN;  // columns [0, lim] are live
      if (causal) lim = min(lim, p.delta + row - c0);
      if (mask_tail && j == last_kblk) lim = min(lim, Nk - 1 - c0);
      step(std::true_type{}, (j & 1) ? tS1 : tS0, bar_s + (j & 1), p_ready + (j & 1), (j >> 1) & 1, j, sc, lim);
    }
    }

    // ---- epilogue ------------------------------------------------------------------------------------
    ptx::mbar_wait(bar_final, 0, 32);
    ptx::tc_fence_after();
    attn_epilogue<D, PV>(p, tOl, l, m_ref, row, Nq, b, hq, orow_base, (PV == PV_E4M3) ? s_vs : nullptr,
                         (PV == PV_E4M3 && p.v_mean) ? s_vm : nullptr);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == HW) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ================================================================================================================
// attn_fwd_n64_kernel -- head_dim 64, INT8 / packed-INT4 K, fp16 P.V: 64-key steps at FOUR CTAs per SM.
//
// 16 softmax warps per SM at 96 registers, 128 TMEM columns per CTA: S [0,64) | O [64,128).  A key block is scored
// as two 32-key HALVES with their own MMAs and barriers: S_lo in columns [0,32), S_hi in [32,64); P_j (32 columns of
// fp16 pairs) goes over S_hi.  A thread streams its row through registers one half at a time and the halves are
// software-pipelined ACROSS key blocks, so that the softmax warps never wait for the tensor pipe in steady state:
//
//   softmax, block j  (S_lo(j) is already in registers)          issuer
//   P_lo = exp2(...)   (kept in 16 registers)
//   wait S_hi(j); ld -> regs;  P_hi = exp2(...)
//   st P_lo, P_hi -> columns [32,64)
//   wait S_lo(j+1); ld -> regs                                     (issued one whole step ago: no wait)
//   arrive p_ready(j)  ------------------------------------------>  PV(j);  QK_hi(j+1) -> columns [32,64) (behind PV(j));
//                                                                   QK_lo(j+2) -> columns [0,32)
//   block j+1: P_lo(j+1) is computed while PV(j) and QK_hi(j+1) run.
//
// (With one S buffer and QK_{j+1} issued whole after PV_j, every softmax warp spent ~23 % of its time waiting for the
// next scores -- ncu source view, profiles/r2_attn_c2_ncu_summary.json.)
//
// Running maximum: a 32-score chunk is first exponentiated against the row's current reference maximum ("optimistic");
// only when its row sum shows a p >= 2^15 (or the chunk is masked, or the row has no reference yet) is the chunk redone
// from its registers with the exact maximum, O and the pending P_lo being rescaled.  Nothing ever re-reads S.
// ================================================================================================================
template <int KM>
struct N64Smem {
  static constexpr int D = 64, BN = 64, VS = 3;
  static constexpr int KS = 3;                                  // int8 operand stages: K_{j+1}, K_{j+2} in use, one refilling
  static constexpr int KPS = 3;                                 // TMA-filled stages (packed INT4: staging ring)
  static constexpr int kQ = kBM * D, kK = BN * D, kKp = BN * D / 2, kV = BN * D * 2;
  static constexpr int kC = 2304;     // constant operand of the bias MMA (see the kernel)
  static constexpr int kKsc = 512;    // k_scale window: 128 blocks
  static constexpr int kBytes = kQ + KS * kK + VS * kV + (KM == KM_I8 ? 0 : KPS * kKp) + kC + kKsc + 256 /*barriers*/ +
                                1024 /*align*/;
};

// fp16 pair * alpha in fp32 (rare path: the pending P_lo when the maximum moves inside the hi chunk)
__device__ __forceinline__ uint32_t scale_f16x2(uint32_t v, float alpha) {
  const __half2 h = *reinterpret_cast<const __half2*>(&v);
  const float2 f = __half22float2(h);
  return ptx::pack_f16x2(f.x * alpha, f.y * alpha);
}

template <int KM, int PF, bool DBG>
__global__ void __launch_bounds__(160, 4)
attn_fwd_n64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  static_assert(KM == KM_I8 || KM == KM_K4, "mixed-width K runs on attn_fwd_kernel");
  using SM = N64Smem<KM>;
  using PC = PvCfg<PV_F16>;
  constexpr int D = 64, BN = 64, HN = 32, KS = SM::KS, KPS = SM::KPS, VS = SM::VS;
  static_assert(KS == 3 && VS == 3, "the issuer's stage arithmetic is written for three stages");
  constexpr bool KX = (KM != KM_I8);
  constexpr uint32_t kTmemCols = 128, kColP = 32, kColO = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Everything in the hot loops is addressed as (sb + compile-time offset) in the 32-bit shared window: one base
  // register instead of a generic -> shared conversion per barrier operation.
  constexpr uint32_t oQ = 0, oK = SM::kQ, oV = oK + KS * SM::kK, oKp = oV + VS * SM::kV;
  constexpr uint32_t oC = oKp + (KX ? KPS * SM::kKp : 0);   // bias-MMA operand
  constexpr uint32_t oKsc = oC + SM::kC;                    // k_scale window
  constexpr uint32_t oBar = oKsc + SM::kKsc;
  constexpr uint32_t oBarQ = oBar;
  constexpr uint32_t oKfull = oBarQ + 8;           // [KPS]
  constexpr uint32_t oKfree = oKfull + 8 * KPS;    // [KPS]
  constexpr uint32_t oVfull = oKfree + 8 * KPS;    // [VS]
  constexpr uint32_t oVfree = oVfull + 8 * VS;     // [VS]
  constexpr uint32_t oSlo = oVfree + 8 * VS;       // QK_lo(j) done: columns [0,32) hold scores          (phase j)
  constexpr uint32_t oShi = oSlo + 8;              // QK_hi(j) done (and with it PV(j-1))               (phase j)
  constexpr uint32_t oLoFree = oShi + 8;           // 4 softmax warps have S_lo(0) in registers          (once)
  constexpr uint32_t oPready = oLoFree + 8;        // 4 softmax warps wrote P_j and hold S_lo(j+1)       (phase j)
  constexpr uint32_t oFinal = oPready + 8;         // last PV done
  constexpr uint32_t oSlot = oFinal + 8;           // TMEM base address
  static_assert(oSlot + 4 <= oBar + 256, "barrier block overflow");
  const uint32_t sb = ptx::pin_u32(ptx::smem_u32(smem));
  uint8_t* sQ = smem + oQ;
  uint8_t* sK = smem + oK;
  uint8_t* sKp = smem + oKp;   // packed INT4 staging ring

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool causal = (p.flags & LOWBIT_ATTN_CAUSAL) != 0;
  int qt, hq, b;
  tile_coords(causal, p.csec, qt, hq, b);
  const int hkv = hq / (p.Hq / p.Hkv);

  const TileView tv = tile_view<D>(p, qt, hq, hkv, b, tid);
  if (tv.done) return;
  const int Nq = tv.Nq, Nk = tv.Nk, nkb = tv.nkb, q_row0 = tv.q_row0, k_row0 = tv.k_row0, tb = tv.tb;
  const int64_t qs_idx = tv.qs_idx, ks_base = tv.ks_base, orow_base = tv.orow_base;

  const bool compat = (p.flags & LOWBIT_ATTN_COMPAT_TAIL) != 0;
  const int nk_lim = compat ? nkb * kScaleBlk : Nk;  // keys [0, nk_lim) take part
  int nblk = (nk_lim + BN - 1) / BN;
  const int dq = p.delta + qt * kBM;  // causal: key c is visible to tile row r iff c <= dq + r
  if (causal) nblk = max(0, min(nblk, (dq + kBM + BN - 1) / BN));
  if (nblk == 0) {
    empty_ring_step<D>(p, qt, hq, b, Nq, tid);
    return;
  }

  if (warp == 4) {
    ptx::tmem_alloc(reinterpret_cast<uint32_t*>(smem + oSlot), kTmemCols);
    ptx::tmem_relinquish();
  }
  if (tid == 0) {
    auto init = [&](uint32_t off, uint32_t count) { ptx::mbar_init(reinterpret_cast<uint64_t*>(smem + off), count); };
    init(oBarQ, 1);
    for (int i = 0; i < KPS; ++i) { init(oKfull + 8 * i, 1); init(oKfree + 8 * i, 1); }
    for (int i = 0; i < VS; ++i) { init(oVfull + 8 * i, 1); init(oVfree + 8 * i, 1); }
    init(oSlo, 1);
    init(oShi, 1);
    init(oLoFree, 4);
    init(oPready, 4);
    init(oFinal, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  // Bias operand: every 16-byte row chunk holds fp16 {1024 x 6, 0 x 2}.  The bias MMA (kind::f16, M 128, N 32, K 16,
  // A = B = this tile, no swizzle, every core matrix 128 bytes after the previous one in both directions) puts
  // 12 * 1024 * 1024 = 12582912.0f = 0x4B400000 into each score column; the int8 MMAs then ACCUMULATE on those bits as
  // int32, so a score leaves TMEM as the fp32 number 12582912 + s and the softmax needs no int -> float conversion
  // (softmax_chunk.cuh, BIASED).  |s| <= 64 * 128 * 128 = 2^20 < 2^22: exact.
  if (tid < SM::kC / 16) ptx::sts_v4_a(sb + oC + tid * 16, 0x64006400u, 0x64006400u, 0x64006400u, 0u);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + oSlot);

  if (warp == 4) {
    // ================================ helper warp ================================
    // one elected lane: TMA producer + tcgen05 issuer.  Packed INT4 K: all 32 lanes also expand the K tiles into the
    // int8 operand stages, three blocks ahead of their QK (K_{j+3} goes into the stage QK_j read -- complete, since
    // S_hi(j) has been consumed).
    const int lane = tid & 31;
    auto run = [&](auto whole_warp_tag) {
      constexpr bool WW = decltype(whole_warp_tag)::value;  // every lane runs this; single-lane work is elected
    constexpr uint32_t idesc_qk = ptx::make_idesc(ptx::kCS32, ptx::kS8, ptx::kS8, 0, 0, kBM, HN);
    constexpr uint32_t idesc_pv = ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 1, kBM, D);
    constexpr uint32_t idesc_bias = ptx::make_idesc(ptx::kCF32, ptx::kF16, ptx::kF16, 0, 0, kBM, HN);
    const uint32_t tS = tmem_base, tP = tmem_base + kColP, tO = tmem_base + kColO;
    const uint64_t dc0 = ptx::make_smem_desc(sb + oC, 128, 128, ptx::kSwzNone);
    // operand descriptors: K-major int8 rows of 64 bytes, 64B swizzle, 8 rows = 512 B (SBO); V: MN-major (d
    // contiguous), 128B swizzle, 8 key rows = 1024 B (SBO), one 64-wide d atom.  Stage / half / k-slice are byte
    // offsets added to the start-address field.
    const uint64_t dq0 = ptx::make_smem_desc(sb + oQ, 16, 8 * D, ptx::kSwz64);
    const uint64_t dk0 = ptx::make_smem_desc(sb + oK, 16, 8 * D, ptx::kSwz64);
    const uint64_t dv0 = ptx::make_smem_desc(sb + oV, BN * 128, 1024, ptx::kSwz128);
    // three-stage rings: tile t lives in stage t % 3 and completes phase (t / 3) & 1 of its barriers; (j3, jd) =
    // (j % 3, j / 3) are carried, not divided
    int j3 = 0, jd = 0;
    auto ring = [&](int off, int& st, uint32_t& par) {  // stage and parity of tile j + off, off in [0, 3]
      int s3 = j3 + off;
      const int w = (s3 >= 3) ? 1 : 0;
      st = s3 - 3 * w;
      par = (uint32_t)(jd + w) & 1u;
    };
    auto load_k_at = [&](int j, int ks, uint32_t par) {  // lead lane: int8 tile (swizzled) into K stage ks (INT8 K only)
      ptx::mbar_wait_a(sb + oKfree + 8 * ks, par ^ 1, 10);
      ptx::mbar_expect_tx_a(sb + oKfull + 8 * ks, SM::kK);
      ptx::tma_load_4d_a(sb + oK + ks * SM::kK, &tmK, sb + oKfull + 8 * ks, 0, k_row0 + j * BN, hkv, tb);
    };
    auto load_kp = [&](int j) {  // lead lane: packed INT4 tile (linear) into staging stage j % 4
      const int ks = j % KPS;
      ptx::mbar_wait_a(sb + oKfree + 8 * ks, ((j / KPS) & 1) ^ 1, 10);
      ptx::mbar_expect_tx_a(sb + oKfull + 8 * ks, SM::kKp);
      ptx::tma_load_4d_a(sb + oKp + ks * SM::kKp, &tmK, sb + oKfull + 8 * ks, 0, k_row0 + j * BN, hkv, tb);
    };
    auto load_v_at = [&](int j, int vs, uint32_t par) {
      ptx::mbar_wait_a(sb + oVfree + 8 * vs, par ^ 1, 11);
      ptx::mbar_expect_tx_a(sb + oVfull + 8 * vs, SM::kV);
      ptx::tma_load_4d_a(sb + oV + vs * SM::kV, &tmV, sb + oVfull + 8 * vs, 0, k_row0 + j * BN, hkv, tb);
    };
    auto expand = [&](int j) {  // whole warp: packed K_j -> operand stage j % KS; frees and refills its staging stage
      const int kps = j % KPS;
      ptx::mbar_wait_a(sb + oKfull + 8 * kps, (j / KPS) & 1, 34);
      unpack_k4_tile<D, BN, 32>(sKp + kps * SM::kKp, sK + (j % KS) * SM::kK, lane);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (!WW || ptx::elect_one()) {
        ptx::mbar_arrive_a(sb + oKfree + 8 * kps);
        if (j + KPS < nblk) load_kp(j + KPS);
      }
    };
    // lead lane: scores of keys [32 half, 32 half + 32) of the tile in K stage ks -> columns [32 half, 32 half + 32)
    auto issue_qk = [&](int ks, uint32_t par, int half) {
      if constexpr (!KX) {
        if (half == 0) ptx::mbar_wait_a(sb + oKfull + 8 * ks, par, 20);
      }
      ptx::tc_fence_after();
      const uint64_t dk = ptx::desc_add(dk0, ks * SM::kK + half * (HN * D));  // 32 rows of 64 bytes = 4 swizzle atoms
      ptx::umma_f16_ss(tS + half * HN, dc0, dc0, idesc_bias, 0);  // every column = 0x4B400000
      ptx::umma_i8_ss(tS + half * HN, dq0, dk, idesc_qk, 1);
      ptx::umma_i8_ss(tS + half * HN, ptx::desc_add(dq0, 32), ptx::desc_add(dk, 32), idesc_qk, 1);
      if (half == 0) {
        ptx::umma_commit_a(sb + oSlo);
      } else {
        ptx::umma_commit_a(sb + oShi);
        if constexpr (!KX) ptx::umma_commit_a(sb + oKfree + 8 * ks);
      }
    };
    if (!WW || ptx::elect_one()) {
      ptx::mbar_expect_tx_a(sb + oBarQ, SM::kQ);
      ptx::tma_load_4d_a(sb + oQ, &tmQ, sb + oBarQ, 0, q_row0 + qt * kBM, hq, tb);
      if constexpr (!KX) {
        for (int j = 0; j < min(KS, nblk); ++j) load_k_at(j, j, 0);
      } else {
        for (int j = 0; j < min(KPS, nblk); ++j) load_kp(j);
      }
      for (int j = 0; j < min(VS - 1, nblk); ++j) load_v_at(j, j, 0);
    }
    if constexpr (KX) {
      __syncwarp();
      ptx::mbar_wait_a(sb + oBarQ, 0, 33);
      permute_q_tile<D, 32>(sQ, lane);
      ptx::fence_proxy_async_smem();
      expand(0);
      if (nblk > 1) expand(1);
      if (nblk > 2) expand(2);
      __syncwarp();
    }
    if (!WW || ptx::elect_one()) {
      if constexpr (!KX) ptx::mbar_wait_a(sb + oBarQ, 0, 21);
      issue_qk(0, 0, 0);
      issue_qk(0, 0, 1);
      if (nblk > 1) {
        ptx::mbar_wait_a(sb + oLoFree, 0, 25);  // every softmax warp holds S_lo(0) in registers
        issue_qk(1, 0, 0);
      }
    }
    for (int j = 0; j < nblk; ++j) {
      if (!WW || ptx::elect_one()) {
        int st;
        uint32_t par;
        ring(0, st, par);
        ptx::mbar_wait_a(sb + oVfull + 8 * st, par, 23);
        ptx::mbar_wait_a(sb + oPready, j & 1, 22);
        ptx::tc_fence_after();
        const uint64_t dv = ptx::desc_add(dv0, st * SM::kV);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          ptx::umma_f16_ts(tO, tP + kk * 8, ptx::desc_add(dv, kk * 16 * 128), idesc_pv, (j > 0) || (kk > 0));
        ptx::umma_commit_a(sb + oVfree + 8 * st);
        if (j == nblk - 1) ptx::umma_commit_a(sb + oFinal);
        if (j + 1 < nblk) {  // overwrites P_j: ordered after PV_j on the tensor pipe
          ring(1, st, par);
          issue_qk(st, par, 1);
        }
        if (j + 2 < nblk) {  // columns [0,32): S_lo(j+1) is in registers (p_ready(j))
          ring(2, st, par);
          issue_qk(st, par, 0);
          load_v_at(j + 2, st, par);  // the stage of PV_{j-1} (complete: QK_hi(j)'s commit followed it)
        }
        if constexpr (!KX) {
          if (j + 3 < nblk) {  // the stage of QK_j (complete: S_hi(j) was consumed)
            ring(3, st, par);
            load_k_at(j + 3, st, par);
          }
        }
        if (++j3 == 3) { j3 = 0; ++jd; }
      }
      if constexpr (KX) {
        __syncwarp();
        if (j + 3 < nblk) expand(j + 3);
        __syncwarp();
      }
    }
    };
    // INT8 K: only the elected lane has work (the other lanes wait at the reconvergence point and cost no issue slots)
    if constexpr (KX) run(std::true_type{});
    else if (ptx::elect_one()) run(std::false_type{});
  } else {
    // ================================ softmax warps ================================
    const int lane = tid & 31;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;  // TMEM lane quadrant of this warp
    const int row = qt * kBM + tid;
    const float qs = p.q_scale[qs_idx] * (KX ? 0.0625f : 1.f);  // packed INT4 K: the operand holds code*16
    const float* ks_ptr = p.k_scale + ks_base;
    const bool mask_tail = (nk_lim % BN) != 0;
    const int last_kblk = (nk_lim + BN - 1) / BN - 1;
    const uint32_t tSl = tmem_base + lane_off, tPl = tSl + kColP, tOl = tSl + kColO;
    constexpr float kScMax = 1.0e-3f;
    float m_ref = -INFINITY, l = 0.f;
    bool have_ref = false;  // warp-uniform: every row of this warp has a finite reference maximum (it never goes back)

    // One 32-score chunk, in registers: pk = fp16 pairs of exp2(s * sc - m_ref), l += their sum.  HI: the chunk is the
    // upper half of block j and plo holds the lower half's P, not yet stored.
    auto chunk_step = [&](auto masked_tag, auto hi_tag, const uint32_t (&s)[32], uint32_t (&pk)[16], uint32_t (&plo)[16],
                          const int j, const float sc, const int lim, const bool fast_ok) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      constexpr bool HI = decltype(hi_tag)::value;
#ifndef LOWBIT_TRACE
      if constexpr (DBG) {
        if (p.dbg != nullptr && j == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
          for (int c = 0; c < 32; ++c) p.dbg[tid * 64 + (HI ? 32 : 0) + c] = (int)(s[c] - chunk::kScoreBias);
        }
      }
#endif
      // The optimistic path reads the biased scores as fp32 (no conversion); the bias leaves through the FMA addend
      // nm - 12582912 * sc, whose rounding puts a common factor of up to 2^(2^-25 * 12582912 * sc) on the chunk's
      // P: <= 1.0004 while sc <= kScMax (typical INT8 / packed-INT4 scales: 1e-4 .. 3e-4, i.e. < 1.0001 -- below the
      // fp16 rounding of P).  Coarser scales (INT4 codes handed over one per int8, tiny head scales) take the exact
      // path, which converts the integer score itself.
      bool exact = !fast_ok;
      if (!exact) {
        // optimistic: keep the reference maximum; the chunk's row sum proves that no p reached 2^15 (finite in fp16).
        // Masked chunks (causal diagonal, tail keys) take it too: a masked column's p is replaced by 0 after the exp2,
        // whatever it was, and stays out of the row sum
        const float lsum = chunk::chunk_f16<MASKED, PF, true>(s, sc, PC::OFF - m_ref, lim, pk);
        exact = __any_sync(0xffffffffu, !(lsum < 32768.f));
        if (!exact) l += lsum;
      }
      if (exact) {
        const int imax = chunk::row_max_i<32, MASKED>(s, lim);  // biased scores keep their order as int32
        const float mblk = (MASKED && imax == INT_MIN) ? -INFINITY : (float)(imax - (int)chunk::kScoreBias) * sc;
        if (__any_sync(0xffffffffu, mblk > m_ref + PC::THR)) {
          const float m_new = fmaxf(m_ref, mblk);
          const float alpha = (m_new == -INFINITY) ? 1.f : ptx::ex2(m_ref - m_new);  // m_ref == -inf -> 0
          l *= alpha;
          m_ref = m_new;
          if constexpr (HI) {
#pragma unroll
            for (int i = 0; i < 16; ++i) plo[i] = scale_f16x2(plo[i], alpha);
          }
          if (j > 0) {
            if constexpr (!HI) {
              ptx::mbar_wait_a(sb + oShi, j & 1, 35);
              ptx::tc_fence_after();
            }
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
        if (!have_ref) have_ref = __all_sync(0xffffffffu, m_ref != -INFINITY);
        const float nm = (m_ref == -INFINITY) ? 0.f : PC::OFF - m_ref;
        l += softmax_block_f16<32, MASKED, 2>(s, sc, nm, lim, pk);  // converts the exact int32 score
      }
    };

    uint32_t s[32];  // on entry to a step: S_lo(j)
#ifdef LOWBIT_TRACE
    const bool tracer = DBG && p.dbg != nullptr && blockIdx.x == 16 && blockIdx.y == 5 && blockIdx.z == 1 && (tid & 31) == 0;
    auto stamp = [&](int j, int k) {
      if (tracer) { int c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); p.dbg[(warp * 64 + j) * 8 + k] = c; }
    };
#else
    auto stamp = [&](int, int) {};
#endif
    auto step = [&](auto masked_tag, const int j, const float sc, const int lim) {
      uint32_t plo[16], phi[16];
      const bool coarse = sc > kScMax;  // warp-uniform (one scale per Q tile and key block)
      stamp(j, 0);
      chunk_step(masked_tag, std::false_type{}, s, plo, plo, j, sc, lim, have_ref && !coarse);
      stamp(j, 1);
      ptx::mbar_wait_a(sb + oShi, j & 1, 31);
      stamp(j, 2);
      ptx::tc_fence_after();
      ptx::tmem_ld_x16(tSl + HN, s);
      ptx::tmem_ld_x16(tSl + HN + 16, s + 16);
      ptx::tmem_wait_ld();
      stamp(j, 3);
      chunk_step(masked_tag, std::true_type{}, s, phi, plo, j, sc, lim - HN, have_ref && !coarse);
      stamp(j, 4);
      ptx::tmem_st_x8(tPl, plo);
      ptx::tmem_st_x8(tPl + 8, plo + 8);
      ptx::tmem_st_x8(tPl + 16, phi);
      ptx::tmem_st_x8(tPl + 24, phi + 8);
      if (j + 1 < nblk) {  // the next block's lower scores: issued a whole step ago
        ptx::mbar_wait_a(sb + oSlo, (j + 1) & 1, 30);
        stamp(j, 5);
        ptx::tc_fence_after();
        ptx::tmem_ld_x16(tSl, s);
        ptx::tmem_ld_x16(tSl + 16, s + 16);
        ptx::tmem_wait_ld();
      }
      stamp(j, 6);
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive_elect_a(sb + oPready);  // P_j is in place and columns [0,32) may take S_lo(j+2)
      stamp(j, 7);
    };
    ptx::mbar_wait_a(sb + oSlo, 0, 30);
    ptx::tc_fence_after();
    ptx::tmem_ld_x16(tSl, s);
    ptx::tmem_ld_x16(tSl + 16, s + 16);
    ptx::tmem_wait_ld();
    ptx::tc_fence_before();
    ptx::mbar_arrive_elect_a(sb + oLoFree);

    int n_full = nblk;  // blocks [0, n_full) need no mask
    if (causal) n_full = max(0, min(n_full, (dq + 1) / BN));
    if (mask_tail) n_full = min(n_full, last_kblk);
    // k_scale: a window of 128 blocks in shared memory, refilled by the 128 softmax threads (a per-step global load
    // sat behind its own spill store); one uniform shared load per step
    for (int j = 0; j < nblk; ++j) {
      if ((j & 127) == 0) {
        if (j > 0) ptx::bar_sync(1, 128);  // every warp is done with the previous window
        ptx::sts_f32_a(sb + oKsc + tid * 4, ks_ptr[min(j + tid, nkb - 1)]);
        ptx::bar_sync(1, 128);
      }
      const float sc = qs * ptx::lds_f32_a(sb + oKsc + (j & 127) * 4);
      if (j < n_full) {
        step(std::false_type{}, j, sc, 0);
      } else {
        const int c0 = j * BN;
        int lim = BN;  // columns [0, lim] are live
        if (causal) lim = min(lim, dq + tid - c0);
        if (mask_tail && j == last_kblk) lim = min(lim, nk_lim - 1 - c0);
        step(std::true_type{}, j, sc, lim);
      }
    }

    // ---- epilogue ------------------------------------------------------------------------------------
    ptx::mbar_wait_a(sb + oFinal, 0, 32);
    ptx::tc_fence_after();
    attn_epilogue<D, PV_F16>(p, tOl, l, m_ref, row, Nq, b, hq, orow_base, nullptr, nullptr);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// o = O_acc / l, lse2 = log2(l) + m  (end of a ring / sequence-parallel pass)
__global__ void attn_finalize_kernel(const float* __restrict__ m, const float* __restrict__ l,
                                     const float* __restrict__ oacc, void* __restrict__ o, float* __restrict__ lse,
                                     int Hq, int Nq, int D, int64_t osb, int64_t osh, int64_t osn, int out_dtype,
                                     int64_t total4) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 output elements
  if (i4 >= total4) return;
  const int64_t rowi = i4 / (D / 4);
  const int c = (int)(i4 % (D / 4)) * 4;
  const int n = (int)(rowi % Nq), h = (int)((rowi / Nq) % Hq), b = (int)(rowi / ((int64_t)Nq * Hq));
  const float inv_l = 1.0f / l[rowi];
  const float4 v = *reinterpret_cast<const float4*>(oacc + rowi * D + c);
  uint2 w;
  if (out_dtype == LOWBIT_F16) {
    w.x = ptx::pack_f16x2(v.x * inv_l, v.y * inv_l);
    w.y = ptx::pack_f16x2(v.z * inv_l, v.w * inv_l);
  } else {
    w.x = ptx::pack_bf16x2(v.x * inv_l, v.y * inv_l);
    w.y = ptx::pack_bf16x2(v.z * inv_l, v.w * inv_l);
  }
  *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(o) + (b * osb + h * osh + (int64_t)n * osn + c) * 2) = w;
  if (lse != nullptr && c == 0) lse[rowi] = ptx::lg2(l[rowi]) + m[rowi];
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

// 4-D tensor map: dims (d0 innermost .. d3), byte strides of dims 1..3, box (box0, box1, 1, 1)
int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, const int64_t (&dim)[4],
             const int64_t (&stride_elems)[3], int box0, int box1, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  LOWBIT_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  LOWBIT_CHECK(((uintptr_t)ptr & 15) == 0, "tensor base address must be 16-byte aligned");
  cuuint64_t dims[4], strides[3];
  for (int i = 0; i < 4; ++i) dims[i] = (cuuint64_t)dim[i];
  for (int i = 0; i < 3; ++i) {
    LOWBIT_CHECK((stride_elems[i] * esize) % 16 == 0, "tensor strides must be multiples of 16 bytes");
    strides[i] = (cuuint64_t)(stride_elems[i] * esize);
    // a size-1 dimension may carry any stride in the caller's tensor; TMA still wants a legal (non-zero, 16B) one
    if (strides[i] == 0) strides[i] = 16;
  }
  cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOWBIT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// (ptr, geometry) -> encoded tensor map.  cuTensorMapEncodeTiled costs 1-2 us per map and an attention call needs 3-5
// of them; serving loops come back with the same buffers (allocator reuse, CUDA-graph replays), so a small
// per-thread cache removes that from the launch path.  A map encodes nothing but its arguments: an entry can never
// be stale, whatever the buffer holds.
struct MapKey {
  const void* ptr;
  int64_t dim[4], str[3];
  int dt, esize, box0, box1, swz;
  bool operator==(const MapKey& o) const {
    if (ptr != o.ptr || dt != o.dt || esize != o.esize || box0 != o.box0 || box1 != o.box1 || swz != o.swz) return false;
    for (int i = 0; i < 4; ++i) if (dim[i] != o.dim[i]) return false;
    for (int i = 0; i < 3; ++i) if (str[i] != o.str[i]) return false;
    return true;
  }
};
static int cached_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, const int64_t (&dim)[4],
                      const int64_t (&stride_elems)[3], int box0, int box1, CUtensorMapSwizzle swz) {
  constexpr int kSlots = 32;
  struct Slot { MapKey key; CUtensorMap map; bool used; };
  thread_local Slot slots[kSlots];
  thread_local int next = 0;
  MapKey k{ptr, {dim[0], dim[1], dim[2], dim[3]}, {stride_elems[0], stride_elems[1], stride_elems[2]},
           (int)dt, esize, box0, box1, (int)swz};
  for (int i = 0; i < kSlots; ++i)
    if (slots[i].used && slots[i].key == k) { *m = slots[i].map; return 0; }
  if (make_map(m, ptr, dt, esize, dim, stride_elems, box0, box1, swz)) return 1;
  slots[next].key = k; slots[next].map = *m; slots[next].used = true;
  next = (next + 1) % kSlots;
  return 0;
}

// The opt-in for > 48 KB of dynamic shared memory is a per-DEVICE function attribute: set it on every launch (it is a
// cheap host-side call) instead of once per process, so that a process driving several GPUs works on all of them.
template <int D, int KM, int PV, bool DBG = false>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                       cudaStream_t st, const CUtensorMap* tk8 = nullptr, const CUtensorMap* tk2 = nullptr) {
  auto kern = attn_fwd_kernel<D, KM, PV, DBG>;
  using SM = AttnSmem<D, KM, PV>;
  LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
  dim3 grid((p.Nq + kBM - 1) / kBM, p.Hq, B);
  kern<<<grid, AttnRoles<D, KM>::kThreads, SM::kBytes, st>>>(tq, tk, tv, tk8 ? *tk8 : tk, tk2 ? *tk2 : tk, p);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

// Development switches (read once): LOWBIT_ATTN_N64=0 sends head_dim 64 back to the 32-key-step kernel;
// LOWBIT_ATTN_PF=n (0..3) sets how many of every 8 score pairs take the FMA-pipe exp2 in the 64-key-step kernel.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
constexpr int kDefaultPF = 2;

template <int KM, int PF, bool DBG = false>
static int launch_n64(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                      cudaStream_t st) {
  auto kern = attn_fwd_n64_kernel<KM, PF, DBG>;
  using SM = N64Smem<KM>;
  LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
  dim3 grid((p.Nq + kBM - 1) / kBM, p.Hq, B);
  kern<<<grid, 160, SM::kBytes, st>>>(tq, tk, tv, p);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}
template <int KM>
static int dispatch_n64(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, int B,
                        cudaStream_t st) {
  static const int pf = env_int("LOWBIT_ATTN_PF", kDefaultPF);
  if (KM == KM_I8 && p.dbg != nullptr) return launch_n64<KM_I8, kDefaultPF, true>(tq, tk, tv, p, B, st);
  switch (pf) {
    case 0: return launch_n64<KM, 0>(tq, tk, tv, p, B, st);
    case 1: return launch_n64<KM, 1>(tq, tk, tv, p, B, st);
    case 3: return launch_n64<KM, 3>(tq, tk, tv, p, B, st);
    default: return launch_n64<KM, 2>(tq, tk, tv, p, B, st);
  }
}
static bool use_n64(int D, int km, int pv, int flags) {
  static const int on = env_int("LOWBIT_ATTN_N64", 1);
  return on != 0 && !(flags & LOWBIT_ATTN_NARROW) && D == 64 && (km == KM_I8 || km == KM_K4) && pv == PV_F16;
}

template <int D>
static int dispatch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                         int B, int km, int pv, cudaStream_t st, const CUtensorMap* tk8, const CUtensorMap* tk2) {
  if (km == KM_MIX) {
    if (pv == PV_F16) return launch_attn<D, KM_MIX, PV_F16>(tq, tk, tv, p, B, st, tk8, tk2);
    return launch_attn<D, KM_MIX, PV_E4M3>(tq, tk, tv, p, B, st, tk8, tk2);
  }
  if (km == KM_F16) return launch_attn<D, KM_F16, PV_F16>(tq, tk, tv, p, B, st);
  if (km == KM_I8 && pv == PV_F16) {
    if (p.dbg != nullptr) return launch_attn<D, KM_I8, PV_F16, true>(tq, tk, tv, p, B, st);
    return launch_attn<D, KM_I8, PV_F16>(tq, tk, tv, p, B, st);
  }
  if (km == KM_K4 && pv == PV_F16) return launch_attn<D, KM_K4, PV_F16>(tq, tk, tv, p, B, st);
  if (km == KM_I8 && pv == PV_E4M3) return launch_attn<D, KM_I8, PV_E4M3>(tq, tk, tv, p, B, st);
  return launch_attn<D, KM_K4, PV_E4M3>(tq, tk, tv, p, B, st);
}

struct AttnArgs {
  const void *q_codes, *k_codes, *v;
  const float *q_scale, *k_scale, *v_scale, *v_mean;
  const int32_t* kbits;
  int B, Hq, Hkv, Nq, Nk, D;
  int64_t qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn;
  int qk_mode, pv_mode, flags;
  // varlen: B = 1, Nq / Nk = packed token counts; the grid covers nseq sequences of at most max_q query rows
  const int32_t *cu_q = nullptr, *cu_k = nullptr, *cu_qs = nullptr, *cu_ks = nullptr;
  int nseq = 0, max_q = 0, nqb_stride = 0, nkb_stride = 0;
};

static int run_attn(const char* who, const AttnArgs& a, AttnParams& p, cudaStream_t st) {
  LOWBIT_CHECK(a.q_codes && a.k_codes && a.v && a.q_scale && (a.k_scale || a.qk_mode == LOWBIT_QK_F16), "%s: null pointer", who);
  LOWBIT_CHECK(a.D == 64 || a.D == 128, "%s: head_dim must be 64 or 128 (got %d)", who, a.D);
  LOWBIT_CHECK(a.B > 0 && a.Hq > 0 && a.Hkv > 0 && a.Nq > 0 && a.Nk > 0, "%s: empty tensor", who);
  LOWBIT_CHECK(a.Hq % a.Hkv == 0, "%s: num_qo_heads (%d) must be divisible by num_kv_heads (%d)", who, a.Hq, a.Hkv);
  LOWBIT_CHECK(a.qk_mode == LOWBIT_QK_I8 || a.qk_mode == LOWBIT_QK_Q8K4 || a.qk_mode == LOWBIT_QK_Q8KMIX ||
                   a.qk_mode == LOWBIT_QK_F16, "%s: bad qk_mode %d", who, a.qk_mode);
  LOWBIT_CHECK(a.qk_mode != LOWBIT_QK_F16 || (a.pv_mode == LOWBIT_PV_F16 && a.cu_q == nullptr),
               "%s: the fp16-QK mode runs with fp16 P.V on padded tensors only", who);
  LOWBIT_CHECK(a.qk_mode != LOWBIT_QK_Q8KMIX || a.kbits != nullptr, "%s: mixed-width K needs kbits", who);
  LOWBIT_CHECK(a.pv_mode == LOWBIT_PV_F16 || a.pv_mode == LOWBIT_PV_E4M3, "%s: bad pv_mode %d", who, a.pv_mode);
  LOWBIT_CHECK(a.pv_mode != LOWBIT_PV_E4M3 || a.v_scale != nullptr, "%s: the FP8 P.V path needs v_scale", who);
  const int D = a.D;
  {
    // causal tile order (tile_coords): as many (batch, head) pairs per longest-first section as keep the section's K / V
    // within ~64 MB of the 126 MB L2 (measured at C2-causal, C3 @ 8K / 16K and D=128 INT8 @ 8K: sections beyond
    // the L2 cost up to 15 %, sections of one head up to 10 %; profiles/r2_causal_order.txt).  LOWBIT_CAUSAL_SECTION overrides.
    static const int forced = [] { const char* e = getenv("LOWBIT_CAUSAL_SECTION"); return e ? atoi(e) : 0; }();
    const double kb = (a.qk_mode == LOWBIT_QK_Q8K4) ? 0.5 : (a.qk_mode == LOWBIT_QK_F16 ? 2.0 : 1.0);
    const double vb = (a.pv_mode == LOWBIT_PV_E4M3) ? 1.0 : 2.0;
    const double per_head = (double)a.Nk * a.D * (kb + vb) / (double)(a.Hq / a.Hkv);
    const double fit = 64.0 * 1024 * 1024 / (per_head > 1.0 ? per_head : 1.0);
    const int nbh = a.B * a.Hq, cap = fit < 1.0 ? 1 : (fit > (double)nbh ? nbh : (int)fit);
    const int nsec = (nbh + cap - 1) / cap;            // equal sections: a short last one would start its heavy tiles
    p.csec = forced > 0 ? forced : (nbh + nsec - 1) / nsec;  // when little else is left to run next to them
  }
  const int km = (a.qk_mode == LOWBIT_QK_Q8K4) ? KM_K4 : (a.qk_mode == LOWBIT_QK_Q8KMIX ? KM_MIX : (a.qk_mode == LOWBIT_QK_F16 ? KM_F16 : KM_I8));
  const int pv = (a.pv_mode == LOWBIT_PV_E4M3) ? PV_E4M3 : PV_F16;
  const bool n64 = use_n64(D, km, pv, a.flags);
  const int BN = n64 ? 64 : ((D == 64) ? AttnCfg<64>::BN : AttnCfg<128>::BN);  // keys per step = rows of a K / V box

  CUtensorMap tq, tk, tv, tk8, tk2;
  const CUtensorMapSwizzle swz_qk = (D == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (km == KM_F16) {  // fp16 / bf16 q, k: 64-element (128-byte) swizzle atoms, like V
    const int64_t dq_[4] = {D, a.Nq, a.Hq, a.B}, sq_[3] = {a.qsn, a.qsh, a.qsb};
    const int64_t dk_[4] = {D, a.Nk, a.Hkv, a.B}, sk_[3] = {a.ksn, a.ksh, a.ksb};
    if (cached_map(&tq, a.q_codes, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dq_, sq_, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (cached_map(&tk, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dk_, sk_, 64, BN, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    const int64_t dim[4] = {D, a.Nq, a.Hq, a.B}, str[3] = {a.qsn, a.qsh, a.qsb};
    if (cached_map(&tq, a.q_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D, kBM, swz_qk)) return 1;
  }
  if (km == KM_F16) {
  } else if (km == KM_I8) {
    const int64_t dim[4] = {D, a.Nk, a.Hkv, a.B}, str[3] = {a.ksn, a.ksh, a.ksb};
    if (cached_map(&tk, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D, BN, swz_qk)) return 1;
  } else if (km == KM_K4) {  // packed INT4: rows of D/2 bytes, landed linearly (no swizzle) for the in-kernel expansion
    const int64_t dim[4] = {D / 2, a.Nk, a.Hkv, a.B}, str[3] = {a.ksn, a.ksh, a.ksb};
    if (cached_map(&tk, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D / 2, BN, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  } else {  // mixed width: container rows of D bytes; one map per bit width, the box takes the row prefix in use
    const int64_t dim[4] = {D, a.Nk, a.Hkv, a.B}, str[3] = {a.ksn, a.ksh, a.ksb};
    if (cached_map(&tk8, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D, BN, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    if (cached_map(&tk, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D / 2, BN, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    if (cached_map(&tk2, a.k_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, D / 4, BN, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  }
  if (pv == PV_F16) {
    const int64_t dim[4] = {D, a.Nk, a.Hkv, a.B}, str[3] = {a.vsn, a.vsh, a.vsb};
    if (cached_map(&tv, a.v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dim, str, 64, BN, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {  // e4m3 [b][h][d][pos]: (vsb, vsh, vsn) are the byte strides of (b, h, d); positions contiguous
    const int64_t npad = (a.Nk + 63) / 64 * 64;
    const int64_t dim[4] = {npad, D, a.Hkv, a.B}, str[3] = {a.vsn, a.vsh, a.vsb};
    if (cached_map(&tv, a.v, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, dim, str, BN, D,
                   BN == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (BN == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)))
      return 1;
  }
  p.q_scale = a.q_scale; p.k_scale = a.k_scale; p.v_scale = a.v_scale; p.v_mean = a.v_mean; p.kbits = a.kbits;
  p.Hq = a.Hq; p.Hkv = a.Hkv; p.Nq = a.Nq; p.Nk = a.Nk;
  p.nqb = (a.Nq + 127) / 128; p.nkb = (a.Nk + 63) / 64;
  p.flags = a.flags;
  int grid_b = a.B;
  if (a.cu_q != nullptr) {  // varlen: the kernel takes per-sequence lengths from cu_*; p.Nq only sizes the grid
    p.cu_q = a.cu_q; p.cu_k = a.cu_k; p.cu_qs = a.cu_qs; p.cu_ks = a.cu_ks;
    p.Nq = a.max_q; p.nqb = a.nqb_stride; p.nkb = a.nkb_stride;
    grid_b = a.nseq;
  }
  const CUtensorMap* p8 = (km == KM_MIX) ? &tk8 : nullptr;
  const CUtensorMap* p2 = (km == KM_MIX) ? &tk2 : nullptr;
  if (n64) {
    if (km == KM_I8) return dispatch_n64<KM_I8>(tq, tk, tv, p, grid_b, st);
    return dispatch_n64<KM_K4>(tq, tk, tv, p, grid_b, st);
  }
  if (D == 64) return dispatch_attn<64>(tq, tk, tv, p, grid_b, km, pv, st, p8, p2);
  return dispatch_attn<128>(tq, tk, tv, p, grid_b, km, pv, st, p8, p2);
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
  LOWBIT_CHECK(o != nullptr, "lowbit_attn_fwd: null pointer");
  LOWBIT_CHECK(out_dtype == LOWBIT_F16 || out_dtype == LOWBIT_BF16, "lowbit_attn_fwd: bad out_dtype %d", out_dtype);
  const bool causal = flags & LOWBIT_ATTN_CAUSAL;
  LOWBIT_CHECK(!causal || Nq == Nk, "lowbit_attn_fwd: causal attention requires qo_len == kv_len");
  LOWBIT_CHECK((osn % 8) == 0 && (osh % 8) == 0 && (osb % 8) == 0 && ((uintptr_t)o & 15) == 0,
               "lowbit_attn_fwd: output must keep 16-byte row alignment");
  AttnArgs a{q_codes, k_codes, v, q_scale, k_scale, v_scale, v_mean, kbits, B, Hq, Hkv, Nq, Nk, D,
             qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn, qk_mode, pv_mode, flags};
  AttnParams p{};
  p.o = o; p.lse = lse; p.osb = osb; p.osh = osh; p.osn = osn;
  p.out_dtype = out_dtype; p.delta = 0; p.first = 1;
  p.dbg = (qk_mode == LOWBIT_QK_I8 && pv_mode == LOWBIT_PV_F16 && !causal) ? g_attn_debug : nullptr;
  return run_attn("lowbit_attn_fwd", a, p, (cudaStream_t)stream);
}

extern "C" int lowbit_attn_fwd_varlen(const void* q_codes, const void* k_codes, const void* v, const float* q_scale,
                                      const float* k_scale, const int32_t* kbits, const int32_t* cu_seqlens_q,
                                      const int32_t* cu_seqlens_k, const int32_t* cu_q_scale, const int32_t* cu_k_scale,
                                      void* o, int nseq, int Hq, int Hkv, int Tq, int Tk, int max_seqlen_q, int D,
                                      int64_t qsh, int64_t qsn, int64_t ksh, int64_t ksn, int64_t vsh, int64_t vsn,
                                      int64_t osh, int64_t osn, int q_scale_stride, int k_scale_stride, int qk_mode,
                                      int out_dtype, int flags, void* stream) {
  LOWBIT_CHECK(o && cu_seqlens_q && cu_seqlens_k && cu_q_scale && cu_k_scale, "lowbit_attn_fwd_varlen: null pointer");
  LOWBIT_CHECK(out_dtype == LOWBIT_F16 || out_dtype == LOWBIT_BF16, "lowbit_attn_fwd_varlen: bad out_dtype %d", out_dtype);
  LOWBIT_CHECK(nseq > 0 && max_seqlen_q > 0, "lowbit_attn_fwd_varlen: empty batch");
  LOWBIT_CHECK(!(flags & LOWBIT_ATTN_COMPAT_TAIL),
               "lowbit_attn_fwd_varlen: compat_tail is not defined for packed sequences (the rows after a sequence are the next one's)");
  LOWBIT_CHECK((osn % 8) == 0 && (osh % 8) == 0 && ((uintptr_t)o & 15) == 0,
               "lowbit_attn_fwd_varlen: output must keep 16-byte row alignment");
  // packed [T,H,D] tensors are one batch entry for the tensor maps; the batch stride is irrelevant
  AttnArgs a{q_codes, k_codes, v, q_scale, k_scale, nullptr, nullptr, kbits, 1, Hq, Hkv, Tq, Tk, D,
             qsn * Tq, qsh, qsn, ksn * Tk, ksh, ksn, vsn * Tk, vsh, vsn, qk_mode, LOWBIT_PV_F16, flags};
  a.cu_q = cu_seqlens_q; a.cu_k = cu_seqlens_k; a.cu_qs = cu_q_scale; a.cu_ks = cu_k_scale;
  a.nseq = nseq; a.max_q = max_seqlen_q; a.nqb_stride = q_scale_stride; a.nkb_stride = k_scale_stride;
  AttnParams p{};
  p.o = o; p.lse = nullptr; p.osb = 0; p.osh = osh; p.osn = osn;
  p.out_dtype = out_dtype; p.delta = 0; p.first = 1;
  return run_attn("lowbit_attn_fwd_varlen", a, p, (cudaStream_t)stream);
}

extern "C" int lowbit_attn_fwd_partial(const void* q_codes, const void* k_codes, const void* v, const float* q_scale,
                                       const float* k_scale, const float* v_scale, const float* v_mean,
                                       const int32_t* kbits, float* m_io, float* l_io, float* o_acc_io,
                                       int B, int Hq, int Hkv, int Nq, int Nk, int D,
                                       int64_t qsb, int64_t qsh, int64_t qsn, int64_t ksb, int64_t ksh, int64_t ksn,
                                       int64_t vsb, int64_t vsh, int64_t vsn, int64_t q_offset, int64_t k_offset,
                                       int qk_mode, int pv_mode, int flags, int first, void* stream) {
  LOWBIT_CHECK(m_io && l_io && o_acc_io, "lowbit_attn_fwd_partial: null state pointer");
  LOWBIT_CHECK(((uintptr_t)o_acc_io & 15) == 0, "lowbit_attn_fwd_partial: o_acc must be 16-byte aligned");
  LOWBIT_CHECK(!(flags & LOWBIT_ATTN_COMPAT_TAIL), "lowbit_attn_fwd_partial: compat_tail is not defined for ring steps");
  int64_t delta = q_offset - k_offset;
  if (delta > (1 << 30)) delta = 1 << 30;
  if (delta < -(1 << 30)) delta = -(1 << 30);
  AttnArgs a{q_codes, k_codes, v, q_scale, k_scale, v_scale, v_mean, kbits, B, Hq, Hkv, Nq, Nk, D,
             qsb, qsh, qsn, ksb, ksh, ksn, vsb, vsh, vsn, qk_mode, pv_mode, flags};
  AttnParams p{};
  p.m_io = m_io; p.l_io = l_io; p.oacc_io = o_acc_io;
  p.delta = (int)delta; p.first = first ? 1 : 0; p.out_dtype = LOWBIT_F16;
  return run_attn("lowbit_attn_fwd_partial", a, p, (cudaStream_t)stream);
}

extern "C" int lowbit_attn_finalize(const float* m, const float* l, const float* o_acc, void* o, float* lse,
                                    int B, int Hq, int Nq, int D, int64_t osb, int64_t osh, int64_t osn,
                                    int out_dtype, void* stream) {
  LOWBIT_CHECK(m && l && o_acc && o, "lowbit_attn_finalize: null pointer");
  LOWBIT_CHECK(D % 4 == 0 && B > 0 && Hq > 0 && Nq > 0, "lowbit_attn_finalize: bad shape");
  LOWBIT_CHECK(out_dtype == LOWBIT_F16 || out_dtype == LOWBIT_BF16, "lowbit_attn_finalize: bad out_dtype %d", out_dtype);
  LOWBIT_CHECK((osn % 4) == 0 && (osh % 4) == 0 && (osb % 4) == 0 && ((uintptr_t)o & 7) == 0,
               "lowbit_attn_finalize: output must keep 8-byte alignment");
  const int64_t total4 = (int64_t)B * Hq * Nq * (D / 4);
  attn_finalize_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      m, l, o_acc, o, lse, Hq, Nq, D, osb, osh, osn, out_dtype, total4);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

