// quant.cu -- memory-bound quantize kernels for sm_100a (HBM3e roofline kernels).
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   quant_per_block_int8_kernel / quant_per_block_int4_unpack_kernel  src/triton/quant_per_block.py:132-178, :22-71
//   QuantInt8Kernel (+ fused mean subtraction)                        csrc/fused/fused.cu:64-198
//   k.mean(dim=seq) and `k - km`                                      src/core.py:293, quant_per_block.py:186-187
//
// Design (B200): one CTA per quantization block; every thread owns 8 contiguous head-dim elements
// (one 128-bit streaming load) of several rows, keeps them in registers as fp32, the block abs-max is a
// redux.sync + one shared-memory hop, and codes leave as 8/4/2-byte packed stores.  The tensor is read
// from HBM exactly once; K smoothing is fused (no materialised `k - km`).  All arithmetic that defines
// the codes uses explicit round-to-nearest intrinsics (__fmul_rn/__fdiv_rn/...) so nvcc can neither
// contract to FMA nor substitute approximate division: codes and scales are bit-exact against the
// IEEE-fp32 restatement of the reference.  This translation unit is compiled WITHOUT --use_fast_math.
#include "common.cuh"
#include "ptx.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <type_traits>

#include <cooperative_groups.h>

namespace lowbit {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}
int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return 1;
}

// ------------------------------------------------------------------------------------------------
// per-block symmetric quantizer
// ------------------------------------------------------------------------------------------------
constexpr int kQuantThreads = 256;

// ---- per-element arithmetic shared by both kernels ------------------------------------------------------
// Q1 code of x for a block scale sc (rcp = RN(1/sc)): y = RN(x / sc) by Markstein's sequence (q0 = RN(x*rcp), exact
// FMA remainder, corrected FMA: the correctly rounded quotient in three instructions where __fdiv_rn takes ~25),
// y += copysign(0.5, y), truncate.  The caller guarantees 1e-30 < sc < 1e30 (so no NaN and no exponent special cases).
__device__ __forceinline__ int q1_code_fast(float x, float sc, float rcp) {
  const float q0 = __fmul_rn(x, rcp);
  const float rem = __fmaf_rn(-q0, sc, x);
  float y = __fmaf_rn(rem, rcp, q0);
  y = __fadd_rn(y, __uint_as_float(0x3f000000u | (__float_as_uint(y) & 0x80000000u)));
  return __float2int_rz(y);
}
// copysign(0.5f, y) in one LOP3: (bits(y) & 0x80000000) | 0x3f000000
__device__ __forceinline__ float half_with_sign_of(float y) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x80000000, 0x3f000000, 0xEA;" : "=r"(r) : "r"(__float_as_uint(y)));
  return __uint_as_float(r);
}
__device__ __forceinline__ int q1_code_ieee(float x, float sc) {
  float y = __fdiv_rn(x, sc);
  y = __fadd_rn(y, y >= 0.f ? 0.5f : -0.5f);
  return (y == y) ? __float2int_rz(y) : 0;  // 0/0 block -> code 0
}
// The reference's kernels as Triton JIT-compiles them for a GPU: fp32 `/` lowers to PTX div.full.f32 (an approximate,
// <= 2 ulp division), for the scale (`max|x| / 127`) and for every quotient (`x / scale`); everything else (mul, add,
// cvt.rzi) is IEEE and uncontracted.  The same instruction here gives the same bits on the same architecture.
__device__ __forceinline__ float div_full(float a, float b) {
  float r;
  asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ int q1_code_divfull(float x, float sc) {
  float y = div_full(x, sc);
  y = __fadd_rn(y, y >= 0.f ? 0.5f : -0.5f);
  return (y == y) ? __float2int_rz(y) : 0;
}
// four int32 codes in [-128,127] -> one word of int8 (cvt.pack: two instructions)
__device__ __forceinline__ uint32_t pack_s8x4(int c0, int c1, int c2, int c3) {
  // d[7:0] = sat(b), d[15:8] = sat(a), d[31:16] = c[15:0]
  uint32_t hi, w;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(c3), "r"(c2));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(c1), "r"(c0), "r"(hi));
  return w;
}
// store 8 codes of one row chunk: int8 (8 B), packed INT4 (4 B) or packed INT2 (2 B)
__device__ __forceinline__ void store_codes8(int8_t* dst_row, int c8, const int (&c)[8], int bits, int pack) {
  if (bits == 8 || !pack) {
    uint2 w;
    w.x = pack_s8x4(c[0], c[1], c[2], c[3]);
    w.y = pack_s8x4(c[4], c[5], c[6], c[7]);
    *reinterpret_cast<uint2*>(dst_row + c8) = w;
  } else if (bits == 4) {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= (uint32_t)(c[i] & 0xf) << (4 * i);
    *reinterpret_cast<uint32_t*>(dst_row + c8 / 2) = w;
  } else {
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= (uint32_t)(c[i] & 0x3) << (2 * i);
    *reinterpret_cast<uint16_t*>(dst_row + c8 / 4) = (uint16_t)w;
  }
}
// the 8 codes of one thread's row chunk under the block's rounding convention
__device__ __forceinline__ void codes8(const float (&x)[8], int (&c)[8], float sc, float rcp, bool triton, bool slow_div,
                                       bool gpu_div) {
  if (triton && gpu_div) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = q1_code_divfull(x[i], sc);
  } else if (triton && !slow_div) {
    // q1_code_fast on packed fp32 pairs (FMUL2 / FFMA2 / FADD2); fma(q0, -sc, x) == fma(-q0, sc, x) bit for bit
    const float2 rcp2 = make_float2(rcp, rcp), nsc2 = make_float2(-sc, -sc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 xv = make_float2(x[2 * i], x[2 * i + 1]);
      const float2 q0 = __fmul2_rn(xv, rcp2);
      const float2 rem = __ffma2_rn(q0, nsc2, xv);
      float2 y = __ffma2_rn(rem, rcp2, q0);
      y = __fadd2_rn(y, make_float2(half_with_sign_of(y.x), half_with_sign_of(y.y)));
      c[2 * i] = __float2int_rz(y.x);
      c[2 * i + 1] = __float2int_rz(y.y);
    }
  } else if (triton) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = q1_code_ieee(x[i], sc);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = max(-128, min(127, __float2int_rn(__fmul_rn(x[i], rcp))));
  }
}

// passes [PB, PB + PC) of x hold the rows of the block that starts at row0 (PC = -1: all passes)
template <int NPASS, int PB = 0, int PC = -1>
__device__ __forceinline__ void quantize_rows(const float (&x)[NPASS][8], int8_t* dst, int64_t osn, int row0, int rpp,
                                              int r0, int c8, int blk, int N, float sc, float rcp, bool triton,
                                              bool slow_div, int bits, int pack, bool gpu_div = false) {
  constexpr int PE = (PC < 0) ? NPASS : PB + PC;
#pragma unroll
  for (int p = PB; p < PE; ++p) {
    const int rl = (p - PB) * rpp + r0;
    const int row = row0 + rl;
    if (!(rl < blk && row < N)) continue;
    int c[8];
    codes8(x[p], c, sc, rcp, triton, slow_div, gpu_div);
    store_codes8(dst + (int64_t)row * osn, c8, c, bits, pack);
  }
}

// One quantization block (jb, h, b) by the whole CTA (kQuantThreads threads).  Shared by the per-tensor kernel below
// and the fused Q+K preparation kernel; ends with every thread past its last read of the shared scratch only after a
// __syncthreads of the NEXT call (callers that loop put a __syncthreads between blocks).
// scale / reciprocal of one block from its abs-max (block-uniform)
struct BlockScale { float sc, rcp; bool slow_div, gpu_div; };
__device__ __forceinline__ BlockScale block_scale(float bmax, int bits, int mode) {
  const float qmax = bits == 8 ? 127.f : (bits == 4 ? 7.f : 1.f);
  BlockScale r;
  r.rcp = 0.f;
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  r.slow_div = (mode & LOWBIT_QMODE_FLAG_IEEE_DIV) != 0;
  r.gpu_div = triton && (mode & LOWBIT_QMODE_FLAG_DIV_FULL) != 0;
  if (r.gpu_div) {
    r.sc = div_full(bmax, qmax);
  } else if (triton) {
    r.sc = __fdiv_rn(bmax, qmax);
    r.rcp = __frcp_rn(r.sc);  // correctly rounded 1/scale, once per block
    // bf16 blocks near the ends of the fp32 exponent range (1/scale denormal or infinite), and all-zero blocks,
    // take the IEEE division: a block-uniform branch
    r.slow_div = r.slow_div || !(r.sc > 1e-30f && r.sc < 1e30f);
  } else {
    bmax = fmaxf(bmax, 1e-7f);
    r.sc = __fdiv_rn(bmax, qmax);
    r.rcp = __fdiv_rn(qmax, bmax);
  }
  return r;
}

// NSUB consecutive quantization blocks of BLK rows, starting at block jb0 of (h, b), by the whole CTA (kQuantThreads
// threads): all rows are loaded first (NSUB * BLK * D * 2 bytes in flight per CTA), then each block gets its own
// abs-max, scale and codes.  Shared by the per-tensor kernel below and the fused Q+K preparation kernel; callers
// that loop put a __syncthreads between calls (s_w: NSUB * kQuantThreads/32 floats of shared scratch).
template <typename T, int D, int BLK, int NSUB = 1>
__device__ __forceinline__ void quant_block_body(const T* __restrict__ in, const T* __restrict__ km,
                                                 int8_t* __restrict__ out, float* __restrict__ scale, int N, int nblk,
                                                 int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh,
                                                 int64_t osn, float sm, int bits, int pack, int mode, int H,
                                                 int jb0, int h, int b, float* s_w) {
  constexpr int TPR = D / 8;                 // threads per row
  constexpr int RPP = kQuantThreads / TPR;   // rows per pass
  constexpr int NPS = (BLK + RPP - 1) / RPP; // passes per block
  constexpr int NP = NSUB * NPS;
  constexpr int NW = kQuantThreads / 32;
  static_assert(NSUB == 1 || BLK % RPP == 0, "sub-blocks must align with passes");
  const int tid = threadIdx.x;
  const int c8 = (tid % TPR) * 8;
  const int r0 = tid / TPR;
  const T* src = in + b * isb + h * ish + c8;
  const int row_base = jb0 * BLK;

  float kmf[8];
  const bool has_km = km != nullptr;
  uint4 km_raw = make_uint4(0, 0, 0, 0);
  if (has_km) {  // L2 load: in the fused kernel km was written by another CTA of the same launch
    km_raw = __ldcg(reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8));
    unpack8<T>(km_raw, kmf);
  }

  float x[NP][8];
  uint4 raw[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int rl = (p % NPS) * RPP + r0;  // row inside its block
    const int row = row_base + (p / NPS) * BLK + rl;
    raw[p] = km_raw;  // rows >= N contribute 0 (masked load): zeros without km, km - km with it
    if (rl < BLK && row < N) raw[p] = ld_stream_v4(src + (int64_t)row * isn);
  }
  float amax[NSUB];
#pragma unroll
  for (int u = 0; u < NSUB; ++u) amax[u] = 0.f;
  auto prep = [&](auto has_km_tag, auto round_tag) {  // block-uniform choice made once, not per element
#pragma unroll
    for (int p = 0; p < NP; ++p)
      prep_row8<T, decltype(has_km_tag)::value, decltype(round_tag)::value>(raw[p], kmf, sm, x[p], amax[p / NPS]);
  };
  if (!has_km) prep(std::false_type{}, std::false_type{});
  else if ((mode & 0xff) == LOWBIT_QMODE_TRITON) prep(std::true_type{}, std::true_type{});  // `k - km` in the input dtype
  else prep(std::true_type{}, std::false_type{});
  // block abs-max: warp redux + smem
#pragma unroll
  for (int u = 0; u < NSUB; ++u) {
    const float m = warp_max(amax[u]);
    if ((tid & 31) == 0) s_w[u * NW + (tid >> 5)] = m;
  }
  __syncthreads();
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  int8_t* dst = out + b * osb + h * osh;
  auto finish = [&](auto u_tag) {
    constexpr int U = decltype(u_tag)::value;
    float bmax = s_w[U * NW];
#pragma unroll
    for (int w = 1; w < NW; ++w) bmax = fmaxf(bmax, s_w[U * NW + w]);
    const BlockScale bs = block_scale(bmax, bits, mode);
    const int jb = jb0 + U;
    if (jb < nblk) {
      if (tid == 0) scale[((int64_t)b * H + h) * nblk + jb] = bs.sc;
      quantize_rows<NP, U * NPS, NPS>(x, dst, osn, jb * BLK, RPP, r0, c8, BLK, N, bs.sc, bs.rcp, triton, bs.slow_div,
                                      bits, pack, bs.gpu_div);
    }
  };
  finish(std::integral_constant<int, 0>{});
  if constexpr (NSUB > 1) finish(std::integral_constant<int, 1>{});
  if constexpr (NSUB > 2) finish(std::integral_constant<int, 2>{});
  if constexpr (NSUB > 3) finish(std::integral_constant<int, 3>{});
}

template <typename T, int D, int BLK>
__global__ void __launch_bounds__(kQuantThreads)
quant_per_block_kernel(const T* __restrict__ in, const T* __restrict__ km, int8_t* __restrict__ out,
                       float* __restrict__ scale, int N, int nblk,
                       int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                       float sm, int bits, int pack, int mode, int H) {
  __shared__ float s_w[kQuantThreads / 32];
  quant_block_body<T, D, BLK>(in, km, out, scale, N, nblk, isb, ish, isn, osb, osh, osn, sm, bits, pack, mode, H,
                              blockIdx.x, blockIdx.y, blockIdx.z, s_w);
}

// The hot configurations of the K chain -- Q1 rounding, int8 or packed INT4 codes, smoothing fused, scale factor 1 -- with those
// arguments as literals: the same body, bit for bit, minus its run-time dispatch (the general kernel spends ~40 % of
// its instructions on mode / width / packing branches, constant-bank loads and row bookkeeping), NSUB blocks per CTA.
template <typename T, int D, int BLK, int NSUB, int BITS, int PACK>
__global__ void __launch_bounds__(kQuantThreads)
quant_k_q1_kernel(const T* __restrict__ in, const T* __restrict__ km, int8_t* __restrict__ out,
                  float* __restrict__ scale, int N, int nblk,
                  int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn, int H) {
  __shared__ float s_w[NSUB * (kQuantThreads / 32)];
  quant_block_body<T, D, BLK, NSUB>(in, km, out, scale, N, nblk, isb, ish, isn, osb, osh, osn, 1.0f, BITS, PACK,
                                    LOWBIT_QMODE_TRITON, H, blockIdx.x * NSUB, blockIdx.y, blockIdx.z, s_w);
}

template <typename T, int D, int BLK>
static int launch_qpb(const void* in, const void* km, void* codes, float* scale, int B, int H, int N,
                      int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                      float sm, int bits, int pack, int mode, cudaStream_t st) {
  const int nblk = (N + BLK - 1) / BLK;
  if constexpr (BLK == 64) {
    static const int fast = [] { const char* e = getenv("LOWBIT_QUANT_FAST"); return e ? atoi(e) : 1; }();
    const bool i8 = bits == 8, i4p = bits == 4 && pack;   // int8 codes (C2) / packed INT4 codes (C3, C4)
    if (fast && km != nullptr && (i8 || i4p) && sm == 1.0f && mode == LOWBIT_QMODE_TRITON) {
      dim3 grid((nblk + 1) / 2, H, B);
      if (i8)
        quant_k_q1_kernel<T, D, BLK, 2, 8, 0><<<grid, kQuantThreads, 0, st>>>(
            (const T*)in, (const T*)km, (int8_t*)codes, scale, N, nblk, isb, ish, isn, osb, osh, osn, H);
      else
        quant_k_q1_kernel<T, D, BLK, 2, 4, 1><<<grid, kQuantThreads, 0, st>>>(
            (const T*)in, (const T*)km, (int8_t*)codes, scale, N, nblk, isb, ish, isn, osb, osh, osn, H);
      LOWBIT_CUDA(cudaGetLastError());
      return 0;
    }
  }
  dim3 grid(nblk, H, B);
  quant_per_block_kernel<T, D, BLK><<<grid, kQuantThreads, 0, st>>>(
      (const T*)in, (const T*)km, (int8_t*)codes, scale, N, nblk, isb, ish, isn, osb, osh, osn, sm, bits, pack, mode, H);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// varlen (packed [T,H,D], cu_seqlens) per-block quantizer: quant_per_block_varlen.py:22-72
// ------------------------------------------------------------------------------------------------
// Grid (ceil(max_seqlen / BLK), H, sequences) like the reference (:120); a CTA whose block lies past its sequence's end
// exits (:44-45).  Blocks restart at every sequence; scales are written head-major, scale[h][cu_scale[b] + jb] with
// `scale_stride` entries per head (the attention kernel's varlen layout), km is one [H,D] row set for all sequences
// (core.py:448: `k.mean(dim=0)`).
template <typename T, int D, int BLK>
__global__ void __launch_bounds__(kQuantThreads)
quant_per_block_varlen_kernel(const T* __restrict__ in, const T* __restrict__ km, int8_t* __restrict__ out,
                              float* __restrict__ scale, const int32_t* __restrict__ cu, const int32_t* __restrict__ cu_scale,
                              int64_t ish, int64_t isn, int64_t osh, int64_t osn, int scale_stride, float sm, int bits,
                              int pack, int mode, int H) {
  __shared__ float s_w[kQuantThreads / 32];
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int start = cu[b], len = cu[b + 1] - start;
  if (jb * BLK >= len) return;
  quant_block_body<T, D, BLK>(in + (int64_t)start * isn, km, out + (int64_t)start * osn, scale + cu_scale[b], len,
                              scale_stride, 0, ish, isn, 0, osh, osn, sm, bits, pack, mode, H, jb, h, 0, s_w);
}

// ------------------------------------------------------------------------------------------------
// mixed-width K: dynamic INT8 / INT4 / INT2 per 64-row block (thresholds of core.py:1055-1061)
// ------------------------------------------------------------------------------------------------
// Container: D bytes per row; a block of width `bits` uses the first D*bits/8 bytes of each of its rows, in the byte
// order the attention kernel's in-smem expansion expects (attn.cu: operand order [d0 d2 d4 d6 d1 d3 d5 d7] inside
// every 8-group):
//   8 bit: byte 8g+p            = code(d = 8g + perm[p]),  perm = [0 2 4 6 1 3 5 7]
//   4 bit: byte 4g+i            = code(8g+2i) | code(8g+2i+1) << 4          (same as the packed INT4 format)
//   2 bit: byte 16c + 8(g%2)+p, bits [2(g/2), 2(g/2)+2) = code(d = 64c + 8g + perm[p]),  g = 8-group inside the
//          64-code chunk c (the kernel's shift k of a chunk yields output bytes 16k .. 16k+15)
// Block statistic = max|k - km| / 127 (compute_scale): > thr8 -> 8 bits, > thr4 -> 4, else 2; or kbits_in when given.
template <typename T, int D>
__global__ void __launch_bounds__(kQuantThreads)
quant_k_mixed_kernel(const T* __restrict__ in, const T* __restrict__ km, const int32_t* __restrict__ kbits_in,
                     int8_t* __restrict__ out, float* __restrict__ scale, int32_t* __restrict__ kbits_out, int N,
                     int nblk, int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                     float thr8, float thr4, int mode, int H) {
  constexpr int BLK = 64;
  constexpr int TPR = D / 8, RPP = kQuantThreads / TPR, NP = BLK / RPP, NW = kQuantThreads / 32;
  __shared__ float s_w[NW];
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* src = in + b * isb + h * ish + c8;
  float kmf[8];
  const bool has_km = km != nullptr;
  if (has_km) unpack8<T>(__ldcg(reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8)), kmf);
  float x[NP][8];
  uint4 raw[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int row = jb * BLK + p * RPP + r0;
    raw[p] = make_uint4(0, 0, 0, 0);
    if (row < N) raw[p] = ld_stream_v4(src + (int64_t)row * isn);
  }
  float amax = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    unpack8<T>(raw[p], x[p]);
    const bool live = jb * BLK + p * RPP + r0 < N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = x[p][i];
      if (has_km) {
        v = __fsub_rn(v, kmf[i]);
        if ((mode & 0xff) == LOWBIT_QMODE_TRITON) v = to_f32<T>(from_f32<T>(v));
      }
      v = live ? v : 0.f;
      x[p][i] = v;
      amax = fmaxf(amax, fabsf(v));
    }
  }
  amax = warp_max(amax);
  if ((tid & 31) == 0) s_w[tid >> 5] = amax;
  __syncthreads();
  float bmax = s_w[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) bmax = fmaxf(bmax, s_w[w]);
  const int64_t bidx = ((int64_t)b * H + h) * nblk + jb;
  int bits;
  if (kbits_in != nullptr) {
    bits = kbits_in[bidx];
  } else {
    const float st = __fdiv_rn(bmax, 127.f);
    bits = st > thr8 ? 8 : (st > thr4 ? 4 : 2);
  }
  const BlockScale bs = block_scale(bmax, bits, mode);
  if (tid == 0) {
    scale[bidx] = bs.sc;
    kbits_out[bidx] = bits;
  }
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  int8_t* dst = out + b * osb + h * osh;
  const int g = (c8 >> 3) & 7, chunk = c8 >> 6;  // 8-group inside its 64-code chunk, chunk inside the row
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int row = jb * BLK + p * RPP + r0;
    const bool live = row < N;
    int c[8];
    codes8(x[p], c, bs.sc, bs.rcp, triton, bs.slow_div, bs.gpu_div);
    int8_t* drow = dst + (int64_t)row * osn;
    if (bits == 8) {
      if (live) {
        uint2 w;
        w.x = pack_s8x4(c[0], c[2], c[4], c[6]);
        w.y = pack_s8x4(c[1], c[3], c[5], c[7]);
        *reinterpret_cast<uint2*>(drow + c8) = w;
      }
    } else if (bits == 4) {
      if (live) store_codes8(drow, c8, c, 4, 1);
    } else {
      // my 8 codes are one 2-bit field (k = g/2) of 8 bytes; the other three fields of those bytes belong to the
      // lanes g^2, g^4, g^6 of the same row: OR them together with two shuffles (block-uniform branch, all lanes in)
      const int sh = 2 * (g >> 1);
      uint32_t lo = 0, hi = 0;
      lo |= (uint32_t)(c[0] & 3) << (sh + 0);  lo |= (uint32_t)(c[2] & 3) << (sh + 8);
      lo |= (uint32_t)(c[4] & 3) << (sh + 16); lo |= (uint32_t)(c[6] & 3) << (sh + 24);
      hi |= (uint32_t)(c[1] & 3) << (sh + 0);  hi |= (uint32_t)(c[3] & 3) << (sh + 8);
      hi |= (uint32_t)(c[5] & 3) << (sh + 16); hi |= (uint32_t)(c[7] & 3) << (sh + 24);
      lo |= __shfl_xor_sync(0xffffffffu, lo, 2); hi |= __shfl_xor_sync(0xffffffffu, hi, 2);
      lo |= __shfl_xor_sync(0xffffffffu, lo, 4); hi |= __shfl_xor_sync(0xffffffffu, hi, 4);
      if (live && (g >> 1) == 0) *reinterpret_cast<uint2*>(drow + chunk * 16 + 8 * (g & 1)) = make_uint2(lo, hi);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-block symmetric quantizer, TMA-pipelined persistent form (the one normally launched)
// ------------------------------------------------------------------------------------------------
// The kernel above lives through load -> reduce -> quantize -> store once per CTA, so the whole chip moves through
// those phases together and HBM idles during the arithmetic.  Here CTAs are persistent (a few per SM), tiles of one
// quantization block ([BLK rows][D] elements, OOB rows zero-filled) arrive through a ring of TMA stages that stays
// several tiles ahead, and the threads quantize tile i while the TMA engine fetches tiles i+1 .. i+STAGES-1.
// Same arithmetic, same results.
template <int D, int BLK> struct QuantTmaCfg {
  static constexpr int kTileBytes = BLK * D * 2;
  static constexpr int kStages = (kTileBytes >= 32768) ? 3 : (kTileBytes >= 16384 ? 4 : 6);
  static constexpr int kSmem = kStages * kTileBytes + 1024 + 256;
};

// HOT: the Q side of the hot path (Q1 rounding, int8 codes, no mean) with those arguments as literals -- the same code
// minus its run-time dispatch, bit-identical by construction.
template <typename T, int D, int BLK, bool HOT = false>
__global__ void __launch_bounds__(kQuantThreads)
quant_per_block_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const T* __restrict__ km,
                           int8_t* __restrict__ out, float* __restrict__ scale, int N, int nblk, int H, int total,
                           int64_t osb, int64_t osh, int64_t osn, float sm, int bits, int pack, int mode) {
  if constexpr (HOT) {
    km = nullptr;
    bits = 8;
    pack = 0;
    mode = LOWBIT_QMODE_TRITON;
  }
  using C = QuantTmaCfg<D, BLK>;
  constexpr int S = C::kStages;
  constexpr int TPR = D / 8;                 // threads per row
  constexpr int RPP = kQuantThreads / TPR;   // rows per pass
  constexpr int NP = (BLK + RPP - 1) / RPP;  // passes
  extern __shared__ uint8_t qsmem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(qsmem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * C::kTileBytes);
  __shared__ float s_w[2][kQuantThreads / 32];
  const int tid = threadIdx.x;
  const int c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  const float qmax = bits == 8 ? 127.f : (bits == 4 ? 7.f : 1.f);

  if (tid == 0) {
    for (int i = 0; i < S; ++i) ptx::mbar_init(full + i, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmIn);
  }
  __syncthreads();
  // Tile t = (b * H + h) * nblk + j.  The coordinates of the tile being quantized and of the tile being fetched
  // advance by a fixed stride per iteration: carried along with two compare-and-wrap steps instead of three integer
  // divisions per tile (which were a sixth of the instructions of this issue-bound kernel).
  struct Coord { int j, h, b; };
  const int first = blockIdx.x, stride = gridDim.x;
  const int dj = stride % nblk, dh = (stride / nblk) % H, db = stride / (nblk * H);
  auto coord_of = [&](int t) { return Coord{t % nblk, (t / nblk) % H, t / (nblk * H)}; };
  auto advance = [&](Coord& c) {
    c.j += dj; c.h += dh; c.b += db;
    if (c.j >= nblk) { c.j -= nblk; c.h += 1; }
    if (c.h >= H) { c.h -= H; c.b += 1; }
  };
  auto issue = [&](const Coord& c, int stage) {  // thread 0: fetch one tile into `stage`
    ptx::mbar_expect_tx(full + stage, C::kTileBytes);
    ptx::tma_load_4d(smem + stage * C::kTileBytes, &tmIn, full + stage, 0, c.j * BLK, c.h, c.b);
  };
  Coord cur = coord_of(first), nxt = cur;  // nxt: the tile S iterations ahead (thread 0's fetch cursor)
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      if (first + i * stride < total) issue(nxt, i);
      advance(nxt);
    }
  }
  int it = 0;
  for (int t = first; t < total; t += stride, ++it, advance(cur)) {
    const int stage = it % S;
    const int jb = cur.j, h = cur.h, b = cur.b;
    float kmf[8];
    const bool has_km = km != nullptr;
    uint4 km_raw = make_uint4(0, 0, 0, 0);
    if (has_km) {
      km_raw = *reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8);
      unpack8<T>(km_raw, kmf);
    }
    ptx::mbar_wait(full + stage, (it / S) & 1, 40);
    const uint8_t* tile = smem + stage * C::kTileBytes;
    float x[NP][8];
    uint4 raw[NP];
    float amax = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int rl = p * RPP + r0;
      const bool live = (rl < BLK) && (jb * BLK + rl < N);
      // rows >= N contribute 0 (the TMA zero fill is the masked load): zeros without km, km - km with it
      raw[p] = km_raw;
      if (live) raw[p] = *reinterpret_cast<const uint4*>(tile + ((size_t)rl * D + c8) * 2);
    }
    auto prep = [&](auto has_km_tag, auto round_tag) {  // block-uniform choice made once, not per element
#pragma unroll
      for (int p = 0; p < NP; ++p)
        prep_row8<T, decltype(has_km_tag)::value, decltype(round_tag)::value>(raw[p], kmf, sm, x[p], amax);
    };
    if (!has_km) prep(std::false_type{}, std::false_type{});
    else if (triton) prep(std::true_type{}, std::true_type{});  // `k - km` in the input dtype
    else prep(std::true_type{}, std::false_type{});
    amax = warp_max(amax);
    if ((tid & 31) == 0) s_w[it & 1][tid >> 5] = amax;
    __syncthreads();  // every thread has finished reading this stage
    if (tid == 0) {
      if (t + S * stride < total) issue(nxt, stage);
      advance(nxt);
    }
    float bmax = s_w[it & 1][0];
#pragma unroll
    for (int w = 1; w < kQuantThreads / 32; ++w) bmax = fmaxf(bmax, s_w[it & 1][w]);

    float sc, rcp = 0.f;
    bool slow_div = (mode & LOWBIT_QMODE_FLAG_IEEE_DIV) != 0;
    const bool gpu_div = triton && (mode & LOWBIT_QMODE_FLAG_DIV_FULL) != 0;
    if (gpu_div) {
      sc = div_full(bmax, qmax);
    } else if (triton) {
      sc = __fdiv_rn(bmax, qmax);
      rcp = __frcp_rn(sc);
      slow_div = slow_div || !(sc > 1e-30f && sc < 1e30f);
    } else {
      bmax = fmaxf(bmax, 1e-7f);
      sc = __fdiv_rn(bmax, qmax);
      rcp = __fdiv_rn(qmax, bmax);
    }
    if (tid == 0) scale[((int64_t)b * H + h) * nblk + jb] = sc;

    int8_t* dst = out + b * osb + h * osh;
    quantize_rows<NP>(x, dst, osn, jb * BLK, RPP, r0, c8, BLK, N, sc, rcp, triton, slow_div, bits, pack, gpu_div);
  }
}

// SM count of the CURRENT device (a process may drive several GPUs: nothing device-specific is cached per process)
static int num_sms() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

template <typename T, int D, int BLK>
static int launch_qpb_tma(const void* in, const void* km, void* codes, float* scale, int B, int H, int N,
                          int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                          float sm, int bits, int pack, int mode, cudaStream_t st) {
  using C = QuantTmaCfg<D, BLK>;
  const int nblk = (N + BLK - 1) / BLK;
  const int64_t total64 = (int64_t)B * H * nblk;
  LOWBIT_CHECK(total64 < (1ll << 31), "lowbit_quant_per_block: too many blocks");
  CUtensorMap tm;
  const int64_t dim[4] = {D, N, H, B}, str[3] = {isn, ish, isb};
  if (make_map(&tm, in, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dim, str, D, BLK, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  static const int fast = [] { const char* e = getenv("LOWBIT_QUANT_FAST"); return e ? atoi(e) : 1; }();
  const bool hot = fast && BLK == 128 && km == nullptr && bits == 8 && mode == LOWBIT_QMODE_TRITON;
  auto kern = hot ? quant_per_block_tma_kernel<T, D, BLK, true> : quant_per_block_tma_kernel<T, D, BLK, false>;
  // the > 48 KB opt-in is a per-device function attribute: set on every launch (cheap), not once per process
  LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
  int ctas_per_sm = (227 * 1024) / (C::kSmem + 1024);
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  const int grid = (int)(total64 < cap ? total64 : cap);
  kern<<<grid, kQuantThreads, C::kSmem, st>>>(tm, (const T*)km, (int8_t*)codes, scale, N, nblk, H, (int)total64,
                                              osb, osh, osn, sm, bits, pack, mode);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K mean: exact, order-independent (fp16) / fp64 (bf16)
// ------------------------------------------------------------------------------------------------
constexpr int kMeanMaxChunks = 16;

__host__ __device__ inline int mean_chunk_rows(int N) {
  int c = (N + kMeanMaxChunks - 1) / kMeanMaxChunks;
  c = (c + 63) / 64 * 64;
  return c < 64 ? 64 : c;
}

template <typename T> struct MeanAcc;
template <> struct MeanAcc<__half> {
  using type = long long;  // units of 2^-24: every finite fp16 is an exact integer
  static __device__ __forceinline__ type cvt(__half v) { return __float2ll_rn(__half2float(v) * 16777216.f); }
  static __device__ __forceinline__ float to_sum_f32(type s) { return __ll2float_rn(s) * 5.9604644775390625e-08f; }
};
template <> struct MeanAcc<__nv_bfloat16> {
  using type = double;
  static __device__ __forceinline__ type cvt(__nv_bfloat16 v) { return (double)__bfloat162float(v); }
  static __device__ __forceinline__ float to_sum_f32(type s) { return __double2float_rn(s); }
};

// partial column sums of one row chunk (ch, h, b) by the whole CTA (256 threads); `s` is [256/(D/8)][D+1] scratch
template <typename T, int D>
__device__ __forceinline__ void ksum_chunk_body(const T* __restrict__ k, typename MeanAcc<T>::type* __restrict__ part,
                                                int N, int chunk, int nchunk, int64_t sb, int64_t sh, int64_t sn,
                                                int H, int ch, int h, int b,
                                                typename MeanAcc<T>::type (*s)[D + 1]) {
  using A = typename MeanAcc<T>::type;
  constexpr int TPR = D / 8, RPP = 256 / TPR;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const T* src = k + b * sb + h * sh + c8;
  const int row_end = min(N, (ch + 1) * chunk);
  A acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = A(0);
  RowAcc<T> ra;
  ra.clear();
  int since_flush = 0;
  // rows in a fixed order inside a thread (matters only for bf16/fp64); 8 loads in flight per thread
  for (int row = ch * chunk + r0; row < row_end; row += 8 * RPP) {
    uint4 ra0[4], ra1[4];
    {
      const T* p = src + (int64_t)row * sn;
      const int64_t st = (int64_t)RPP * sn;
      ld_stream_v4_x4(p, p + st, p + 2 * st, p + 3 * st, row < row_end, row + RPP < row_end, row + 2 * RPP < row_end,
                      row + 3 * RPP < row_end, ra0);
      ld_stream_v4_x4(p + 4 * st, p + 5 * st, p + 6 * st, p + 7 * st, row + 4 * RPP < row_end, row + 5 * RPP < row_end,
                      row + 6 * RPP < row_end, row + 7 * RPP < row_end, ra1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) ra.add(ra0[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) ra.add(ra1[u]);
    since_flush += 8;
    if ((since_flush & 15) == 0) ra.fold();
    if (since_flush >= 256) { ra.flush(acc); since_flush = 0; }
  }
  ra.flush(acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[r0][c8 + i] = acc[i];
  __syncthreads();
  if (tid < D) {
    A t = A(0);
    for (int r = 0; r < RPP; ++r) t += s[r][tid];
    part[(((int64_t)b * H + h) * nchunk + ch) * D + tid] = t;
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(256, 2)
k_mean_partial_kernel(const T* __restrict__ k, typename MeanAcc<T>::type* __restrict__ part, int N, int chunk,
                      int nchunk, int64_t sb, int64_t sh, int64_t sn, int H) {
  using A = typename MeanAcc<T>::type;
  __shared__ A s[256 / (D / 8)][D + 1];
  ksum_chunk_body<T, D>(k, part, N, chunk, nchunk, sb, sh, sn, H, blockIdx.x, blockIdx.y, blockIdx.z, s);
}

template <typename T, int D>
__global__ void k_mean_final_kernel(const typename MeanAcc<T>::type* __restrict__ part, T* __restrict__ km,
                                    int N, int nchunk, int total) {
  using A = typename MeanAcc<T>::type;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over B*H*D
  if (idx >= total) return;
  const int d = idx % D;
  const int64_t bh = idx / D;
  A t = A(0);
  for (int c = 0; c < nchunk; ++c) t += part[(bh * nchunk + c) * D + d];
  const float s32 = MeanAcc<T>::to_sum_f32(t);
  km[idx] = from_f32<T>(__fdiv_rn(s32, (float)N));
}

// ------------------------------------------------------------------------------------------------
// global abs-max (compute_scale, core.py:1039-1047) and LSE fix-up (core.py:296-304,347)
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256)
abs_max_kernel(const T* __restrict__ x, unsigned int* __restrict__ out, int N, int64_t sb, int64_t sh, int64_t sn) {
  constexpr int TPR = D / 8, RPP = 256 / TPR;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const T* src = x + blockIdx.z * sb + blockIdx.y * sh + c8;
  float m = 0.f;
  for (int row = blockIdx.x * RPP * 8 + r0; row < min(N, (int)(blockIdx.x + 1) * RPP * 8); row += RPP) {
    float f[8];
    unpack8<T>(ld_stream_v4(src + (int64_t)row * sn), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(f[i]));
  }
  m = warp_max(m);
  if ((tid & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// global max and min (asymmetric compute_scale, core.py:1043-1045): floats ordered as unsigned integers
// (negative -> ~bits, non-negative -> bits | sign bit), one atomicMax / atomicMin per warp
__device__ __forceinline__ unsigned int float_ordered(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <typename T, int D>
__global__ void __launch_bounds__(256)
min_max_kernel(const T* __restrict__ x, unsigned int* __restrict__ out, int N, int64_t sb, int64_t sh, int64_t sn) {
  constexpr int TPR = D / 8, RPP = 256 / TPR;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const T* src = x + blockIdx.z * sb + blockIdx.y * sh + c8;
  float mx = -INFINITY, mn = INFINITY;
  for (int row = blockIdx.x * RPP * 8 + r0; row < min(N, (int)(blockIdx.x + 1) * RPP * 8); row += RPP) {
    float f[8];
    unpack8<T>(ld_stream_v4(src + (int64_t)row * sn), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { mx = fmaxf(mx, f[i]); mn = fminf(mn, f[i]); }
  }
  const unsigned int omx = __reduce_max_sync(0xffffffffu, float_ordered(mx));
  const unsigned int omn = __reduce_min_sync(0xffffffffu, float_ordered(mn));
  if ((tid & 31) == 0) {
    atomicMax(out, omx);
    atomicMin(out + 1, omn);
  }
}
__global__ void min_max_finish_kernel(unsigned int* __restrict__ out) {  // ordered integers -> float bits, in place
  const unsigned int u = out[threadIdx.x];
  out[threadIdx.x] = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
}

template <typename T>
__global__ void lse_fixup_kernel(float* __restrict__ lse, const T* __restrict__ q, const T* __restrict__ km,
                                 int Hq, int Hkv, int Nq, int D, int64_t sb, int64_t sh, int64_t sn, float sm_scale) {
  // one warp per (b, h, n) row: lse = lse2 / log2e + (q . km) * sm_scale
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int64_t total = (int64_t)gridDim.y * Hq * Nq;
  (void)total;
  const int b = blockIdx.y;
  if (warp >= Hq * Nq) return;
  const int h = warp / Nq, n = warp % Nq;
  float dot = 0.f;
  if (km != nullptr) {
    const T* qr = q + b * sb + h * sh + (int64_t)n * sn;
    const T* kr = km + ((int64_t)b * Hkv + h / (Hq / Hkv)) * D;
    for (int d = lane; d < D; d += 32) dot = fmaf(to_f32<T>(qr[d]), to_f32<T>(kr[d]), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    // the reference computes q.km with a matmul in the input dtype, then casts to fp32 (core.py:296-304)
    dot = to_f32<T>(from_f32<T>(dot));
  }
  if (lane == 0) {
    float* p = lse + ((int64_t)b * Hq + h) * Nq + n;
    *p = __fdiv_rn(*p, 1.44269504f) + dot * sm_scale;
  }
}


// ------------------------------------------------------------------------------------------------
// K smoothing + K codes in ONE launch with ONE pass over HBM (src/core.py:291-306 + quant_per_block.py:181-248)
// ------------------------------------------------------------------------------------------------
// The separate path runs three dependent launches (column-sum chunks -> mean -> quantizer) and reads K twice.  Here a
// thread-block cluster of kKsqCluster CTAs owns one (batch, kv-head) slice: every CTA parks its share of the rows in
// shared memory while it adds them up (the same exact, order-independent fixed-point sum as k_mean_partial_kernel),
// the per-CTA column sums meet through distributed shared memory, every CTA forms the same mean (bit-identical to
// lowbit_k_mean: exact sum -> fp32 -> / N -> fp16), and the 64-row blocks are then quantized straight out of shared
// memory with the arithmetic of quant_block_body.  fp16 only (the bf16 mean is defined by a summation order).
constexpr int kKsqCluster = 8;                 // portable cluster size
constexpr int kKsqMaxTileBytes = 200 * 1024;   // rows of one CTA held in shared memory

__host__ __device__ inline int ksq_rows_per_cta(int N) {
  const int nblk = (N + 63) / 64;
  return (nblk + kKsqCluster - 1) / kKsqCluster * 64;
}

template <typename T, int D>
__global__ void __launch_bounds__(kQuantThreads, (D == 64 ? 3 : 2))
k_smooth_quant_cluster_kernel(const T* __restrict__ k, T* __restrict__ km_out, int8_t* __restrict__ out,
                              float* __restrict__ scale, int N, int nblk, int rows_cta, int64_t isb, int64_t ish,
                              int64_t isn, int64_t osb, int64_t osh, int64_t osn, int bits, int pack, int mode, int H) {
  namespace cg = cooperative_groups;
  using A = typename MeanAcc<T>::type;
  constexpr int TPR = D / 8, RPP = kQuantThreads / TPR, NW = kQuantThreads / 32;
  constexpr int NPS = 64 / RPP;  // passes per 64-row block
  extern __shared__ uint8_t ksq_smem_raw[];
  uint4* tile = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(ksq_smem_raw) + 15) & ~uintptr_t(15));
  __shared__ A s_part[NW][D];      // per-warp column sums
  __shared__ A s_sum[D];           // this CTA's column sums (read by the whole cluster)
  __shared__ __align__(16) T s_km[D];
  __shared__ float s_w[2][NW];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int slice = blockIdx.x / kKsqCluster, h = slice % H, b = slice / H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cq = tid % TPR, c8 = cq * 8, r0 = tid / TPR;
  const int row_begin = rank * rows_cta;
  const T* src = k + b * isb + h * ish + c8;

  // (a) rows -> shared memory, exact column sums on the way
  {
    A acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = A(0);
    RowAcc<T> ra;
    ra.clear();
    int since_flush = 0;
    // every row of this CTA's share is requested at once with 16-byte asynchronous copies straight into shared
    // memory (no registers held while the ~64 KB per CTA are in flight); a thread then sums the chunks it copied
    // itself, so its own wait_group is all the ordering the sum needs
    for (int r = r0; r < rows_cta; r += RPP) {
      if (row_begin + r < N) {
        const uint32_t daddr = (uint32_t)__cvta_generic_to_shared(&tile[r * TPR + cq]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(daddr), "l"(src + (int64_t)(row_begin + r) * isn) : "memory");
      } else {
        tile[r * TPR + cq] = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int r = r0; r < rows_cta; r += RPP) {
      ra.add(tile[r * TPR + cq]);  // absent rows add exact zeros
      ++since_flush;
      if ((since_flush & 15) == 0) ra.fold();
      if (since_flush >= 240) { ra.flush(acc); since_flush = 0; }
    }
    ra.flush(acc);
    // lanes that share lane % TPR hold the same columns: fold them, then one row of sums per warp
#pragma unroll
    for (int off = TPR; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    }
    if (lane < TPR) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_part[warp][c8 + i] = acc[i];
    }
  }
  __syncthreads();
  if (tid < D) {
    A t = A(0);
#pragma unroll
    for (int w = 0; w < NW; ++w) t += s_part[w][tid];
    s_sum[tid] = t;
  }
  // (b) the cluster's column sums -> the slice mean (every CTA computes the same value)
  cluster.sync();
  if (tid < D) {
    A t = A(0);
    for (int rr = 0; rr < kKsqCluster; ++rr) t += cluster.map_shared_rank(s_sum, rr)[tid];
    const T kmv = from_f32<T>(__fdiv_rn(MeanAcc<T>::to_sum_f32(t), (float)N));
    s_km[tid] = kmv;
    if (rank == 0) km_out[((int64_t)b * H + h) * D + tid] = kmv;
  }
  cluster.sync();  // s_km visible to the CTA; nobody leaves while its s_sum may still be read

  // (c) 64-row blocks out of shared memory
  const uint4 km_raw = *reinterpret_cast<const uint4*>(&s_km[c8]);
  float kmf[8];
  unpack8<T>(km_raw, kmf);
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  int8_t* dst = out + b * osb + h * osh;
  const int nloc = rows_cta / 64;
  for (int bl = 0; bl < nloc; ++bl) {
    const int jb = rank * nloc + bl;
    if (jb >= nblk) break;  // uniform over the CTA
    float x[NPS][8];
    uint4 raw[NPS];
    float amax = 0.f;
#pragma unroll
    for (int p = 0; p < NPS; ++p) {
      const int rl = p * RPP + r0;
      raw[p] = km_raw;  // rows >= N contribute 0 (masked load): km - km
      if (jb * 64 + rl < N) raw[p] = tile[(bl * 64 + rl) * TPR + cq];
    }
    if (triton) {
#pragma unroll
      for (int p = 0; p < NPS; ++p) prep_row8<T, true, true>(raw[p], kmf, 1.0f, x[p], amax);
    } else {
#pragma unroll
      for (int p = 0; p < NPS; ++p) prep_row8<T, true, false>(raw[p], kmf, 1.0f, x[p], amax);
    }
    amax = warp_max(amax);
    if (lane == 0) s_w[bl & 1][warp] = amax;
    __syncthreads();
    float bmax = s_w[bl & 1][0];
#pragma unroll
    for (int w = 1; w < NW; ++w) bmax = fmaxf(bmax, s_w[bl & 1][w]);
    const BlockScale bs = block_scale(bmax, bits, mode);
    if (tid == 0) scale[((int64_t)b * H + h) * nblk + jb] = bs.sc;
    quantize_rows<NPS>(x, dst, osn, jb * 64, RPP, r0, c8, 64, N, bs.sc, bs.rcp, triton, bs.slow_div, bits, pack,
                       bs.gpu_div);
  }
}

template <typename T, int D>
static int launch_ksq(const void* k, void* km_out, void* codes, float* scale, int B, int H, int N, int64_t isb,
                      int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn, int bits, int pack, int mode,
                      cudaStream_t st) {
  const int rows_cta = ksq_rows_per_cta(N);
  const int smem = rows_cta * D * 2 + 16;
  auto kern = k_smooth_quant_cluster_kernel<T, D>;
  LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kKsqMaxTileBytes + 16));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((int64_t)B * H * kKsqCluster), 1, 1);
  cfg.blockDim = dim3(kQuantThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kKsqCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int nblk = (N + 63) / 64;
  LOWBIT_CUDA(cudaLaunchKernelEx(&cfg, kern, (const T*)k, (T*)km_out, (int8_t*)codes, scale, N, nblk, rows_cta, isb, ish,
                                 isn, osb, osh, osn, bits, pack, mode, H));
  return 0;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" {

int lowbit_version(void) { return LOWBIT_ABI_VERSION; }
const char* lowbit_last_error(void) { return last_error().c_str(); }

int lowbit_quant_per_block_varlen(const void* in, const void* km, void* codes, float* scale, const int32_t* cu_seqlens,
                                  const int32_t* cu_scale, int nseq, int H, int max_seqlen, int D, int64_t ish,
                                  int64_t isn, int64_t osh, int64_t osn, int scale_stride, int blk, int bits, int pack,
                                  float sm, int mode, int dtype, void* stream) {
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_quant_per_block_varlen: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(in && codes && scale && cu_seqlens && cu_scale, "lowbit_quant_per_block_varlen: null pointer");
  LOWBIT_CHECK(nseq > 0 && H > 0 && max_seqlen > 0, "lowbit_quant_per_block_varlen: empty batch");
  LOWBIT_CHECK(blk == 64 || blk == 128, "lowbit_quant_per_block_varlen: block size must be 64 or 128 (got %d)", blk);
  LOWBIT_CHECK(bits == 8 || bits == 4 || bits == 2, "lowbit_quant_per_block_varlen: bits must be 8, 4 or 2 (got %d)", bits);
  LOWBIT_CHECK((mode & 0xff) == LOWBIT_QMODE_TRITON || (mode & 0xff) == LOWBIT_QMODE_CUDA, "lowbit_quant_per_block_varlen: bad mode %d", mode);
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && ((uintptr_t)in & 15) == 0,
               "lowbit_quant_per_block_varlen: input base address and strides must keep 16-byte alignment");
  const int div = pack ? 8 / bits : 1;
  LOWBIT_CHECK((osn * div) % 8 == 0, "lowbit_quant_per_block_varlen: code rows must keep 8-byte alignment");
  dim3 grid((max_seqlen + blk - 1) / blk, H, nseq);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_VL(T, DD, BB)                                                                                        \
  quant_per_block_varlen_kernel<T, DD, BB><<<grid, kQuantThreads, 0, st>>>((const T*)in, (const T*)km, (int8_t*)codes, \
      scale, cu_seqlens, cu_scale, ish, isn, osh, osn, scale_stride, sm, bits, pack, mode, H)
#define LAUNCH_VL_T(T)                                                                      \
  do {                                                                                      \
    if (D == 64 && blk == 64) LAUNCH_VL(T, 64, 64);                                         \
    else if (D == 64) LAUNCH_VL(T, 64, 128);                                                \
    else if (blk == 64) LAUNCH_VL(T, 128, 64);                                              \
    else LAUNCH_VL(T, 128, 128);                                                            \
  } while (0)
  if (dtype == LOWBIT_F16) LAUNCH_VL_T(__half);
  else if (dtype == LOWBIT_BF16) LAUNCH_VL_T(__nv_bfloat16);
  else return lowbit::fail("lowbit_quant_per_block_varlen: bad dtype %d", dtype);
#undef LAUNCH_VL_T
#undef LAUNCH_VL
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int lowbit_quant_k_mixed(const void* k, const void* km, const int32_t* kbits_in, void* codes, float* scale,
                         int32_t* kbits_out, int B, int H, int N, int D, int64_t isb, int64_t ish, int64_t isn,
                         int64_t osb, int64_t osh, int64_t osn, float thr8, float thr4, int mode, int dtype,
                         void* stream) {
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_quant_k_mixed: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(k && codes && scale && kbits_out, "lowbit_quant_k_mixed: null pointer");
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_quant_k_mixed: empty tensor");
  LOWBIT_CHECK((mode & 0xff) == LOWBIT_QMODE_TRITON || (mode & 0xff) == LOWBIT_QMODE_CUDA, "lowbit_quant_k_mixed: bad mode %d", mode);
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && isb % 8 == 0 && ((uintptr_t)k & 15) == 0,
               "lowbit_quant_k_mixed: input base address and strides must keep 16-byte alignment");
  LOWBIT_CHECK(osn % 16 == 0 && osh % 16 == 0 && osb % 16 == 0 && ((uintptr_t)codes & 15) == 0,
               "lowbit_quant_k_mixed: container rows must keep 16-byte alignment");
  const int nblk = (N + 63) / 64;
  dim3 grid(nblk, H, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_MIX(T, DD)                                                                                         \
  quant_k_mixed_kernel<T, DD><<<grid, kQuantThreads, 0, st>>>((const T*)k, (const T*)km, kbits_in, (int8_t*)codes, \
      scale, kbits_out, N, nblk, isb, ish, isn, osb, osh, osn, thr8, thr4, mode, H)
  if (dtype == LOWBIT_F16) { if (D == 64) LAUNCH_MIX(__half, 64); else LAUNCH_MIX(__half, 128); }
  else if (dtype == LOWBIT_BF16) { if (D == 64) LAUNCH_MIX(__nv_bfloat16, 64); else LAUNCH_MIX(__nv_bfloat16, 128); }
  else return lowbit::fail("lowbit_quant_k_mixed: bad dtype %d", dtype);
#undef LAUNCH_MIX
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int64_t lowbit_k_mean_workspace_bytes(int B, int H, int N, int D) {
  const int chunk = mean_chunk_rows(N);
  const int nchunk = (N + chunk - 1) / chunk;
  return (int64_t)B * H * nchunk * D * 8;
}

int lowbit_k_mean(const void* k, void* km_out, void* workspace, int B, int H, int N, int D,
                  int64_t sb, int64_t sh, int64_t sn, int dtype, void* stream) {
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_k_mean: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(k && km_out && workspace, "lowbit_k_mean: null pointer");
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_k_mean: empty tensor");
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0 && ((uintptr_t)k & 15) == 0,
               "lowbit_k_mean: base address and strides must keep 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = mean_chunk_rows(N);
  const int nchunk = (N + chunk - 1) / chunk;
  dim3 grid(nchunk, H, B);
  const int total = B * H * D;
#define LAUNCH(T, DD)                                                                                          \
  k_mean_partial_kernel<T, DD><<<grid, 256, 0, st>>>((const T*)k, (MeanAcc<T>::type*)workspace, N, chunk,      \
                                                     nchunk, sb, sh, sn, H);                                   \
  k_mean_final_kernel<T, DD><<<(total + 255) / 256, 256, 0, st>>>((const MeanAcc<T>::type*)workspace,          \
                                                                  (T*)km_out, N, nchunk, total);
  if (dtype == LOWBIT_F16) {
    if (D == 64) { LAUNCH(__half, 64) } else { LAUNCH(__half, 128) }
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) { LAUNCH(__nv_bfloat16, 64) } else { LAUNCH(__nv_bfloat16, 128) }
  } else {
    return fail("lowbit_k_mean: unsupported dtype %d", dtype);
  }
#undef LAUNCH
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}


int lowbit_k_smooth_quant_supported(int N, int D, int dtype) {
  if (dtype != LOWBIT_F16 || (D != 64 && D != 128) || N <= 0) return 0;
  return (int64_t)ksq_rows_per_cta(N) * D * 2 <= kKsqMaxTileBytes ? 1 : 0;
}

int lowbit_k_smooth_quant(const void* k, void* km_out, void* codes, float* scale, int B, int H, int N, int D,
                          int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                          int bits, int pack, int mode, int dtype, void* stream) {
  LOWBIT_CHECK(k && km_out && codes && scale, "lowbit_k_smooth_quant: null pointer");
  LOWBIT_CHECK(lowbit_k_smooth_quant_supported(N, D, dtype),
               "lowbit_k_smooth_quant: unsupported (N=%d, D=%d, dtype=%d): fp16, head_dim 64/128, a (b,h) slice must fit "
               "the cluster's shared memory -- use lowbit_k_mean + lowbit_quant_per_block", N, D, dtype);
  LOWBIT_CHECK(bits == 8 || bits == 4 || bits == 2, "lowbit_k_smooth_quant: bits must be 8, 4 or 2 (got %d)", bits);
  LOWBIT_CHECK((mode & 0xff) == LOWBIT_QMODE_TRITON || (mode & 0xff) == LOWBIT_QMODE_CUDA, "lowbit_k_smooth_quant: bad mode %d", mode);
  LOWBIT_CHECK(B > 0 && H > 0 && (int64_t)B * H * kKsqCluster < (1ll << 31), "lowbit_k_smooth_quant: bad batch / head count");
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && isb % 8 == 0 && ((uintptr_t)k & 15) == 0,
               "lowbit_k_smooth_quant: input must keep 16-byte alignment");
  const int ob = (pack && bits < 8) ? bits : 8;  // bytes of 8 codes
  LOWBIT_CHECK(osn % ob == 0 && osh % ob == 0 && osb % ob == 0, "lowbit_k_smooth_quant: output strides misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 64) return launch_ksq<__half, 64>(k, km_out, codes, scale, B, H, N, isb, ish, isn, osb, osh, osn, bits, pack, mode, st);
  return launch_ksq<__half, 128>(k, km_out, codes, scale, B, H, N, isb, ish, isn, osb, osh, osn, bits, pack, mode, st);
}

int lowbit_quant_per_block(const void* in, const void* km, void* codes, float* scale, int B, int H, int N, int D,
                           int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                           int blk, int bits, int pack, float sm, int mode, int dtype, void* stream) {
  LOWBIT_CHECK(in && codes && scale, "lowbit_quant_per_block: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_quant_per_block: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(blk == 32 || blk == 64 || blk == 128, "lowbit_quant_per_block: blk must be 32, 64 or 128 (got %d)", blk);
  LOWBIT_CHECK(bits == 8 || bits == 4 || bits == 2, "lowbit_quant_per_block: bits must be 8, 4 or 2 (got %d)", bits);
  LOWBIT_CHECK((mode & 0xff) == LOWBIT_QMODE_TRITON || (mode & 0xff) == LOWBIT_QMODE_CUDA, "lowbit_quant_per_block: bad mode %d", mode);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_quant_per_block: empty tensor");
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && isb % 8 == 0 && ((uintptr_t)in & 15) == 0,
               "lowbit_quant_per_block: input base address and strides must keep 16-byte alignment");
  const int ob = (pack && bits < 8) ? bits : 8;  // bytes of 8 codes
  LOWBIT_CHECK(osn % ob == 0 && osh % ob == 0 && osb % ob == 0, "lowbit_quant_per_block: output strides misaligned");
  cudaStream_t st = (cudaStream_t)stream;
#define ARGS in, km, codes, scale, B, H, N, isb, ish, isn, osb, osh, osn, sm, bits, pack, mode, st
  // 128-row blocks go through the TMA-pipelined persistent kernel, 64-row blocks through the one-block-per-CTA
  // kernel (measured faster at that tile size: 8-16 KB tiles do not amortise the per-tile barrier).
  // LOWBIT_QUANT_TMA=0 / 2 force neither / both (A/B measurements).
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("LOWBIT_QUANT_TMA"); use_tma = e ? atoi(e) : 1; }
  const bool tma_ok = use_tma != 0;
#define BY_BLK(T, DD)                                                                        \
  switch (blk) {                                                                             \
    case 32: return launch_qpb<T, DD, 32>(ARGS);                                             \
    case 64: return (tma_ok && use_tma > 1) ? launch_qpb_tma<T, DD, 64>(ARGS) : launch_qpb<T, DD, 64>(ARGS); \
    default: return tma_ok ? launch_qpb_tma<T, DD, 128>(ARGS) : launch_qpb<T, DD, 128>(ARGS); \
  }
  if (dtype == LOWBIT_F16) {
    if (D == 64) { BY_BLK(__half, 64) } else { BY_BLK(__half, 128) }
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) { BY_BLK(__nv_bfloat16, 64) } else { BY_BLK(__nv_bfloat16, 128) }
  }
#undef BY_BLK
#undef ARGS
  return fail("lowbit_quant_per_block: unsupported dtype %d", dtype);
}

int lowbit_abs_max(const void* x, float* out, int B, int H, int N, int D, int64_t sb, int64_t sh, int64_t sn,
                   int dtype, void* stream) {
  LOWBIT_CHECK(x && out, "lowbit_abs_max: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_abs_max: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0 && ((uintptr_t)x & 15) == 0,
               "lowbit_abs_max: base address and strides must keep 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  LOWBIT_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  const int rows_per_cta = (256 / (D / 8)) * 8;
  dim3 grid((N + rows_per_cta - 1) / rows_per_cta, H, B);
  if (dtype == LOWBIT_F16) {
    if (D == 64) abs_max_kernel<__half, 64><<<grid, 256, 0, st>>>((const __half*)x, (unsigned*)out, N, sb, sh, sn);
    else abs_max_kernel<__half, 128><<<grid, 256, 0, st>>>((const __half*)x, (unsigned*)out, N, sb, sh, sn);
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) abs_max_kernel<__nv_bfloat16, 64><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (unsigned*)out, N, sb, sh, sn);
    else abs_max_kernel<__nv_bfloat16, 128><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (unsigned*)out, N, sb, sh, sn);
  } else {
    return fail("lowbit_abs_max: unsupported dtype %d", dtype);
  }
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int lowbit_min_max(const void* x, float* out, int B, int H, int N, int D, int64_t sb, int64_t sh, int64_t sn,
                   int dtype, void* stream) {
  LOWBIT_CHECK(x && out, "lowbit_min_max: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_min_max: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_min_max: empty tensor");
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0 && ((uintptr_t)x & 15) == 0,
               "lowbit_min_max: base address and strides must keep 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  LOWBIT_CUDA(cudaMemsetAsync(out, 0x00, sizeof(float), st));        // ordered minimum: below every float
  LOWBIT_CUDA(cudaMemsetAsync(out + 1, 0xff, sizeof(float), st));    // ordered maximum: above every float
  const int rows_per_cta = (256 / (D / 8)) * 8;
  dim3 grid((N + rows_per_cta - 1) / rows_per_cta, H, B);
  if (dtype == LOWBIT_F16) {
    if (D == 64) min_max_kernel<__half, 64><<<grid, 256, 0, st>>>((const __half*)x, (unsigned*)out, N, sb, sh, sn);
    else min_max_kernel<__half, 128><<<grid, 256, 0, st>>>((const __half*)x, (unsigned*)out, N, sb, sh, sn);
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) min_max_kernel<__nv_bfloat16, 64><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (unsigned*)out, N, sb, sh, sn);
    else min_max_kernel<__nv_bfloat16, 128><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (unsigned*)out, N, sb, sh, sn);
  } else {
    return fail("lowbit_min_max: unsupported dtype %d", dtype);
  }
  min_max_finish_kernel<<<1, 2, 0, st>>>((unsigned*)out);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int lowbit_lse_fixup(float* lse, const void* q, const void* km, int B, int Hq, int Hkv, int Nq, int D,
                     int64_t sb, int64_t sh, int64_t sn, float sm_scale, int dtype, void* stream) {
  LOWBIT_CHECK(lse && q, "lowbit_lse_fixup: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t warps = (int64_t)Hq * Nq;
  dim3 grid((unsigned)((warps * 32 + 255) / 256), B);
  if (dtype == LOWBIT_F16)
    lse_fixup_kernel<__half><<<grid, 256, 0, st>>>(lse, (const __half*)q, (const __half*)km, Hq, Hkv, Nq, D, sb, sh, sn, sm_scale);
  else if (dtype == LOWBIT_BF16)
    lse_fixup_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(lse, (const __nv_bfloat16*)q, (const __nv_bfloat16*)km, Hq, Hkv, Nq, D, sb, sh, sn, sm_scale);
  else
    return fail("lowbit_lse_fixup: unsupported dtype %d", dtype);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
