// softmax_chunk.cuh -- register-level softmax building blocks of the head_dim-64 attention kernel (csrc/attn.cu,
// attn_fwd_n64_kernel) and of the instruction-mix microbenchmark tools/ubench_softmax.cu (same code, no TMEM).
//
// One thread owns one query row.  A step hands it 64 int32 scores S (exact QK^T of the int8 codes) as two chunks of
// 32:  p = exp2(S * sc + nm)  ->  fp16 pairs (the A operand of the P.V MMA) + fp32 row sum.
//
// exp2 runs on two pipes at once.  MUFU.EX2 does 16 results / clk / SM and is the binding unit at head_dim 64
// (DESIGN.md 4.2); PF of every 8 score pairs therefore take the FMA pipe instead: Cody-Waite range reduction with
// the 1.5 * 2^23 magic constant, a degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, below the
// 4.9e-4 half-ulp of the fp16 P it is rounded to) and an integer multiply-add that inserts the exponent.  The
// argument is clamped to [-112, 16] by the saturating form of the scaling FMA itself (no extra instruction): 2^-112
// is zero for every purpose of an fp16 P, and 2^16 trips the kernel's overflow check on the row sum (a block whose
// scores outgrow the row's reference maximum by 2^15 is redone with the exact maximum).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace lowbit {
namespace chunk {

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// FMA-pipe exp2 of the pair (sf0, sf1) * sc + nm, arguments clamped to [-112, 16].
//   y = sat(sf * sc/128 + (nm + 112)/128)            in [0, 1]:  x = 128 y - 112
//   r = 128 y + (M - 112)                            M = 1.5 * 2^23: the low mantissa bits of r hold n = round(x)
//   f = 128 y - (r - (M - 112))                      x - n, in [-0.5, 0.5]
//   2^x = bits(poly(f)) + (n << 23)                  (r << 23 drops r's own exponent field: same thing)
struct PolyConst {
  float sc128, nm128;
};
__device__ __forceinline__ PolyConst poly_const(float sc, float nm) {
  PolyConst c;
  c.sc128 = sc * 0.0078125f;
  c.nm128 = fmaf(nm, 0.0078125f, 0.875f);
  return c;
}
__device__ __forceinline__ float2 ex2_poly_pair(float sf0, float sf1, const PolyConst& c) {
  const float kC1 = 12582912.0f - 112.0f;
  const float2 y = make_float2(fma_sat(sf0, c.sc128, c.nm128), fma_sat(sf1, c.sc128, c.nm128));
  const float2 k128 = make_float2(128.f, 128.f);
  const float2 r = __ffma2_rn(y, k128, make_float2(kC1, kC1));
  const float2 t = __fadd2_rn(r, make_float2(-kC1, -kC1));                 // n + 112, exact
  const float2 f = __ffma2_rn(y, k128, make_float2(-t.x, -t.y));          // x - n
  float2 p = __ffma2_rn(make_float2(0.05517103523015976f, 0.05517103523015976f), f,
                        make_float2(0.24260984361171722f, 0.24260984361171722f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999281764030457f, 0.9999281764030457f));
  float2 o;
  o.x = __int_as_float(__float_as_int(r.x) * 0x00800000 + __float_as_int(p.x));
  o.y = __int_as_float(__float_as_int(r.y) * 0x00800000 + __float_as_int(p.y));
  return o;
}

// N (32, or 16) scores -> N/2 packed fp16 pairs + their fp32 sum.  Columns > lim contribute 0 when MASKED.
// PF in [0, 8]: pairs (c/2) % 8 < PF go through the FMA-pipe exp2.
// BIASED: the scores arrive as int32 bit patterns kScoreBias + s (the tensor core accumulated the int8 products on
// top of an fp32 1.5 * 2^23): read as fp32 they ARE 12582912 + s, exactly, so no int -> float conversion is needed and
// the constant leaves through the addend of the scaling FMA (one rounding of nm - 12582912 * sc, |error| <= half an
// ulp of a number of a few thousand: ~1e-4 in the exponent, the same for every score of the chunk and its row sum).
constexpr uint32_t kScoreBias = 0x4B400000u;   // bits of 12582912.0f = 1.5 * 2^23
constexpr float kScoreBiasF = 12582912.0f;
template <bool MASKED, int PF, bool BIASED = false, int N = 32>
__device__ __forceinline__ float chunk_f16(const uint32_t* __restrict__ s, float sc, float nm_in, int lim,
                                           uint32_t* __restrict__ pk) {
  const float nm = BIASED ? fmaf(-kScoreBiasF, sc, nm_in) : nm_in;
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  PolyConst pc = poly_const(sc, nm_in);
  if (BIASED) pc.nm128 = fmaf(-kScoreBiasF, pc.sc128, pc.nm128);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < N; c += 4) {
    float2 p[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = c + 2 * h;
      const float f0 = BIASED ? __uint_as_float(s[cc]) : __int2float_rn((int)s[cc]);
      const float f1 = BIASED ? __uint_as_float(s[cc + 1]) : __int2float_rn((int)s[cc + 1]);
      if (((cc / 2) % 8) < PF) {
        p[h] = ex2_poly_pair(f0, f1, pc);
      } else {
        const float2 x = __ffma2_rn(make_float2(f0, f1), sc2, nm2);
        p[h] = make_float2(ex2_mufu(x.x), ex2_mufu(x.y));
      }
      if (MASKED) {
        p[h].x = (cc <= lim) ? p[h].x : 0.f;
        p[h].y = (cc + 1 <= lim) ? p[h].y : 0.f;
      }
    }
    acc0 = __fadd2_rn(acc0, p[0]);
    acc1 = __fadd2_rn(acc1, p[1]);
    pk[c / 2] = pack_f16x2(p[0].x, p[0].y);
    pk[c / 2 + 1] = pack_f16x2(p[1].x, p[1].y);
  }
  const float2 t = __fadd2_rn(acc0, acc1);
  return t.x + t.y;
}

// integer row maximum of N scores (columns > lim excluded when MASKED): 4 independent chains
template <int N, bool MASKED>
__device__ __forceinline__ int row_max_i(const uint32_t* __restrict__ s, int lim) {
  int m4[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
  for (int c = 0; c < N; ++c) m4[c & 3] = max(m4[c & 3], (!MASKED || c <= lim) ? (int)s[c] : INT_MIN);
  return max(max(m4[0], m4[1]), max(m4[2], m4[3]));
}

}  // namespace chunk
}  // namespace lowbit
