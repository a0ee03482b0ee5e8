// quant_extra.cu -- the finer-grained / packed / FP8 quantizers of the low-bit attention path (HBM-bound kernels).
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   Q3  quant_{query,key}_per_thread_int{8,4}_kernel     src/triton/quant_per_thread.py:22-219
//   Q5  _minmax_along_last_dim + _pack_along_last_dim     src/triton/utils/quant/new_pack.py:198-300
//   Q6  TransposePadPermuteKernel + MeanScaleKernel       csrc/fused/fused.cu:263-428 (src/quant.py:210-291)
// Compiled WITHOUT --use_fast_math: codes, scales and zero points are bit-exact against IEEE restatements.
#include "common.cuh"

#include <cuda_fp8.h>

namespace lowbit {

// ------------------------------------------------------------------------------------------------
// Q3: per-thread-group quantizer.  Q: block of 32 rows, group t = rows {8i+t}; K: block of 64 rows,
// group t = rows {8i+2t, 8i+2t+1}.  scale = amax/QMAX + 1e-7 (quant_per_thread.py:62,114).
// ------------------------------------------------------------------------------------------------
template <typename T, int D, int BLK, bool IS_KEY>
__global__ void __launch_bounds__(256)
quant_per_thread_kernel(const T* __restrict__ in, const T* __restrict__ km, int8_t* __restrict__ out,
                        float* __restrict__ scale, int N, int n_scale, int64_t isb, int64_t ish, int64_t isn,
                        int64_t osb, int64_t osh, int64_t osn, int bits, int H) {
  constexpr int TPR = D / 8, RPP = 256 / TPR, NP = (BLK + RPP - 1) / RPP;
  constexpr int NG = IS_KEY ? 4 : 8;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* src = in + b * isb + h * ish + c8;
  __shared__ unsigned int s_g[NG];
  if (tid < NG) s_g[tid] = 0u;
  float kmf[8];
  const bool has_km = km != nullptr;
  if (has_km) unpack8<T>(*reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8), kmf);
  float x[NP][8];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int rl = p * RPP + r0, row = jb * BLK + rl;
    uint4 raw = make_uint4(0, 0, 0, 0);
    const bool live = rl < BLK && row < N;
    if (live) raw = ld_stream_v4(src + (int64_t)row * isn);
    unpack8<T>(raw, x[p]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = x[p][i];
      if (has_km) v = to_f32<T>(from_f32<T>(__fsub_rn(v, kmf[i])));  // `k - km` in the input dtype (:234-235)
      x[p][i] = live ? v : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int rl = p * RPP + r0;
    if (rl >= BLK) continue;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a = fmaxf(a, fabsf(x[p][i]));
    const int g = IS_KEY ? ((rl & 7) >> 1) : (rl & 7);
    atomicMax(&s_g[g], __float_as_uint(a));
  }
  __syncthreads();
  const float qmax = bits == 8 ? 127.f : 7.f;
  if (tid < NG) {
    const int slot = jb * NG + tid;
    if (slot < n_scale) scale[((int64_t)b * H + h) * n_scale + slot] = __fadd_rn(__fdiv_rn(__uint_as_float(s_g[tid]), qmax), 1e-7f);
  }
  int8_t* dst = out + b * osb + h * osh;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int rl = p * RPP + r0, row = jb * BLK + rl;
    if (!(rl < BLK && row < N)) continue;
    const int g = IS_KEY ? ((rl & 7) >> 1) : (rl & 7);
    const float sc = __fadd_rn(__fdiv_rn(__uint_as_float(s_g[g]), qmax), 1e-7f);
    int c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = __fdiv_rn(x[p][i], sc);
      y = __fadd_rn(y, y >= 0.f ? 0.5f : -0.5f);
      c[i] = __float2int_rz(y);
    }
    uint2 w;
    w.x = (c[0] & 0xff) | ((c[1] & 0xff) << 8) | ((c[2] & 0xff) << 16) | ((c[3] & 0xff) << 24);
    w.y = (c[4] & 0xff) | ((c[5] & 0xff) << 8) | ((c[6] & 0xff) << 16) | ((c[7] & 0xff) << 24);
    *reinterpret_cast<uint2*>(dst + (int64_t)row * osn + c8) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// Q5: KIVI asymmetric group quantize + pack, group = 32 along the last dim, all math in fp16
// (new_pack.py:262-295).  One thread per group: 64 B in, bits*4 B codes + scale + zero point out.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

template <int BITS>
__global__ void __launch_bounds__(256)
kivi_pack_kernel(const __half* __restrict__ data, uint8_t* __restrict__ code, __half* __restrict__ scale,
                 __half* __restrict__ mn_out, int64_t n_groups) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint4* src = reinterpret_cast<const uint4*>(data + g * 32);
  float x[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 raw = src[i];
    const __half* hv = reinterpret_cast<const __half*>(&raw);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[i * 8 + k] = __half2float(hv[k]);
  }
  float mx = x[0], mn = x[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) { mx = fmaxf(mx, x[i]); mn = fminf(mn, x[i]); }
  constexpr float qm = (float)((1 << BITS) - 1);
  const float sc = h_round(__fdiv_rn(h_round(__fsub_rn(mx, mn)), qm));  // fp16: (mx-mn) then / qm
  scale[g] = __float2half_rn(sc);
  mn_out[g] = __float2half_rn(mn);
  uint32_t w[BITS];  // 32 codes * BITS bits = BITS words
#pragma unroll
  for (int i = 0; i < BITS; ++i) w[i] = 0u;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float y = h_round(__fsub_rn(x[i], mn));
    y = h_round(__fdiv_rn(y, sc));
    y = fminf(fmaxf(y, 0.f), qm);              // NaN (0/0, constant group) -> 0
    const int c = (y == y) ? __float2int_rz(__fadd_rn(y, 0.5f)) : 0;  // round half away (y >= 0)
    w[(i * BITS) / 32] |= (uint32_t)c << ((i * BITS) % 32);
  }
  uint32_t* dst = reinterpret_cast<uint32_t*>(code + g * (4 * BITS));
#pragma unroll
  for (int i = 0; i < BITS; ++i) dst[i] = w[i];
}

// ------------------------------------------------------------------------------------------------
// Q6: V -> e4m3, per channel, transposed + padded to 64 + token-permuted inside 16-groups
// ------------------------------------------------------------------------------------------------
template <typename T> struct SumAcc;
template <> struct SumAcc<__half> {
  using type = long long;  // exact fixed point, units of 2^-24
  static __device__ __forceinline__ float to_f32(type s) { return __ll2float_rn(s) * 5.9604644775390625e-08f; }
};
template <> struct SumAcc<__nv_bfloat16> {
  using type = double;
  static __device__ __forceinline__ float to_f32(type s) { return __double2float_rn(s); }
};
constexpr int kVChunks = 16;
struct VPartial { float mx, mn; long long sum_bits; };  // sum as raw 8 bytes (int64 or double)

// pass 1: per-channel max / min (/ exact sum when the mean is wanted) over one chunk of rows; 8 loads in flight
template <typename T, int D, bool SUM>
__global__ void __launch_bounds__(256)
v_stats_partial_kernel(const T* __restrict__ v, VPartial* __restrict__ part, int N, int chunk, int nchunk,
                       int64_t sb, int64_t sh, int64_t sn, int H) {
  using A = typename SumAcc<T>::type;
  constexpr int TPR = D / 8, RPP = 256 / TPR;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const int ch = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* src = v + b * sb + h * sh + c8;
  const int row_end = min(N, (ch + 1) * chunk);
  float mx[8], mn[8];
  A acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mx[i] = -INFINITY; mn[i] = INFINITY; acc[i] = A(0); }
  RowAcc<T> ra;
  ra.clear();
  int since_flush = 0;
  for (int row = ch * chunk + r0; row < row_end; row += 8 * RPP) {
    uint4 raw[8];
    bool live[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      live[u] = row + u * RPP < row_end;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (live[u]) raw[u] = ld_stream_v4(src + (int64_t)(row + u * RPP) * sn);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (!live[u]) continue;
      float f[8];
      unpack8<T>(raw[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { mx[i] = fmaxf(mx[i], f[i]); mn[i] = fminf(mn[i], f[i]); }
      if (SUM) ra.add(raw[u]);
    }
    if (SUM) {
      since_flush += 8;
      if ((since_flush & 15) == 0) ra.fold();
      if (since_flush >= 256) { ra.flush(acc); since_flush = 0; }
    }
  }
  if (SUM) ra.flush(acc);
  __shared__ float s_mx[RPP][D + 1], s_mn[RPP][D + 1];
  __shared__ A s_sum[SUM ? RPP : 1][D + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_mx[r0][c8 + i] = mx[i];
    s_mn[r0][c8 + i] = mn[i];
    if (SUM) s_sum[r0][c8 + i] = acc[i];
  }
  __syncthreads();
  if (tid < D) {
    float a = -INFINITY, c = INFINITY;
    A t = A(0);
    for (int r = 0; r < RPP; ++r) {
      a = fmaxf(a, s_mx[r][tid]);
      c = fminf(c, s_mn[r][tid]);
      if (SUM) t += s_sum[r][tid];
    }
    VPartial o;
    o.mx = a; o.mn = c;
    o.sum_bits = *reinterpret_cast<long long*>(&t);
    part[(((int64_t)b * H + h) * nchunk + ch) * D + tid] = o;
  }
}

// pass 1b: one thread per (b, h, channel): fold the chunk statistics into (1/scale, mean) once, instead of once per
// pass-2 CTA (a 16-deep dependent load chain in front of a 16 KB tile)
template <typename T, int D>
__global__ void v_stats_final_kernel(const VPartial* __restrict__ part, float2* __restrict__ tab,
                                     float* __restrict__ v_scale, float* __restrict__ vm_out, int N, int nchunk,
                                     float scale_max, int total) {
  using A = typename SumAcc<T>::type;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over B*H*D
  if (idx >= total) return;
  const int d = idx % D;
  const int64_t bh = idx / D;
  const int n16 = (N + 15) / 16 * 16;
  float mx = -INFINITY, mn = INFINITY;
  A t = A(0);
  for (int c = 0; c < nchunk; ++c) {
    const VPartial q = part[(bh * nchunk + c) * D + d];
    mx = fmaxf(mx, q.mx); mn = fminf(mn, q.mn);
    t += *reinterpret_cast<const A*>(&q.sum_bits);
  }
  if (n16 > N) { mx = fmaxf(mx, 0.f); mn = fminf(mn, 0.f); }  // statistics run over the zero-padded 16-multiple (:336-337)
  float vm = 0.f, amax;
  if (vm_out != nullptr) {
    vm = __fdiv_rn(SumAcc<T>::to_f32(t), (float)n16);         // mean divides by the padded count (:382)
    amax = fmaxf(fabsf(__fsub_rn(mx, vm)), fabsf(__fsub_rn(mn, vm)));
    vm_out[idx] = vm;
  } else {
    amax = fmaxf(fabsf(mx), fabsf(mn));
  }
  v_scale[idx] = __fdiv_rn(amax, scale_max);
  tab[idx] = make_float2(amax > 0.f ? __fdiv_rn(scale_max, amax) : 0.f, vm);  // all-zero channel: codes 0 instead of 0/0
}

// pass 2: every row-group of TPR threads owns one quad of OUTPUT positions, i.e. the four tokens
// {2w, 2w+1, 8+2w, 9+2w} of a 16-group (fused.cu:290-292 inverted), loads their rows, scales, packs the four e4m3
// codes of every channel into one word, and the CTA writes [channel][position] rows through a padded shared-memory
// transpose.  A CTA takes kVTiles consecutive tiles and issues all their row loads first (8 x 16 B in flight/thread).
constexpr int kVTiles = 2;
template <int D> struct VQuantCfg {
  static constexpr int TPR = D / 8, G = 256 / TPR, TT = G * 4;  // tokens per tile: 128 (D=64) / 64 (D=128)
};
template <typename T, int D>
__global__ void __launch_bounds__(256)
v_fp8_quant_kernel(const T* __restrict__ v, const float2* __restrict__ tab, uint8_t* __restrict__ v8, int N, int npad,
                   int64_t sb, int64_t sh, int64_t sn, int64_t osb, int64_t osh, int64_t osd, int H) {
  using C = VQuantCfg<D>;
  constexpr int TPR = C::TPR, TT = C::TT, WPR = TT / 4;  // words per channel row of a tile
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, g = tid / TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  __shared__ float s_r[D], s_vm[D];
  __shared__ uint32_t s_t[kVTiles][D][WPR + 1];
  const T* src = v + b * sb + h * sh + c8;
  uint4 raw[kVTiles][4];
#pragma unroll
  for (int tt = 0; tt < kVTiles; ++tt) {
    const int tok0 = (blockIdx.x * kVTiles + tt) * TT + (g / 4) * 16 + 2 * (g % 4);
    const int toks[4] = {tok0, tok0 + 1, tok0 + 8, tok0 + 9};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      raw[tt][j] = make_uint4(0, 0, 0, 0);
      if (toks[j] < N) raw[tt][j] = ld_stream_v4(src + (int64_t)toks[j] * sn);
    }
  }
  if (tid < D) {
    const float2 t = tab[((int64_t)b * H + h) * D + tid];
    s_r[tid] = t.x;
    s_vm[tid] = t.y;
  }
  __syncthreads();
#pragma unroll
  for (int tt = 0; tt < kVTiles; ++tt) {
    float f[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack8<T>(raw[tt][j], f[j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float vm = s_vm[c8 + i], r = s_r[c8 + i];
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = __fmul_rn(__fsub_rn(f[j][i], vm), r);
      uint16_t lo, hi;
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(y[1]), "f"(y[0]));
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(y[3]), "f"(y[2]));
      s_t[tt][c8 + i][g] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
  }
  __syncthreads();
  // channel rows of the tiles: kVTiles * TT contiguous output bytes each, one word per thread (coalesced)
  uint8_t* dst = v8 + b * osb + h * osh + (int64_t)blockIdx.x * kVTiles * TT;
  for (int idx = tid; idx < D * kVTiles * WPR; idx += 256) {
    const int d = idx / (kVTiles * WPR), w = idx % (kVTiles * WPR);
    if ((int64_t)blockIdx.x * kVTiles * TT + 4 * w < npad)
      *reinterpret_cast<uint32_t*>(dst + (int64_t)d * osd + 4 * w) = s_t[w / WPR][d][w % WPR];
  }
}


// ------------------------------------------------------------------------------------------------
// sub_mean (src/quant.py:175-207 -> SubMeanKernel, csrc/fused/fused.cu:200-261): v_smoothed = fp16(v - vm), the
// difference taken IN THE INPUT DTYPE (__hsub2 on half2 / bfloat162, :243) and then converted to fp16 (:245-248).
// One thread per 8 channels of one token; HBM-bound: 2 B read + 2 B written per element.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
sub_mean_kernel(const T* __restrict__ v, const T* __restrict__ vm, __half* __restrict__ out, int H, int N, int D,
                int64_t sb, int64_t sh, int64_t sn, int64_t osb, int64_t osh, int64_t osn, int64_t total8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int d8 = D / 8;
  const int c = (int)(i % d8) * 8;
  const int64_t t = i / d8;
  const int n = (int)(t % N), h = (int)((t / N) % H), b = (int)(t / ((int64_t)N * H));
  const uint4 raw = ld_stream_v4(v + b * sb + h * sh + (int64_t)n * sn + c);
  const uint4 mraw = *reinterpret_cast<const uint4*>(vm + ((int64_t)b * H + h) * D + c);
  float x[8], m[8];
  unpack8<T>(raw, x);
  unpack8<T>(mraw, m);
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // fp32 difference of two 16-bit values rounded once to T == the T-precision subtraction of the reference
    const float2 dlt = round_trip2<T>(make_float2(x[2 * k] - m[2 * k], x[2 * k + 1] - m[2 * k + 1]));
    const __half2 hh = __float22half2_rn(dlt);
    ow[k] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  *reinterpret_cast<uint4*>(out + b * osb + h * osh + (int64_t)n * osn + c) = o;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" {

int lowbit_quant_per_thread(const void* in, const void* km, void* codes, float* scale, int B, int H, int N, int D,
                            int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                            int warp_blk, int n_scale, int is_key, int bits, int dtype, void* stream) {
  LOWBIT_CHECK(in && codes && scale, "lowbit_quant_per_thread: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_quant_per_thread: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(bits == 8 || bits == 4, "lowbit_quant_per_thread: bits must be 8 or 4 (got %d)", bits);
  LOWBIT_CHECK((is_key && warp_blk == 64) || (!is_key && warp_blk == 32),
               "lowbit_quant_per_thread: WARPQ must be 32 and WARPK 64 (got %d)", warp_blk);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_quant_per_thread: empty tensor");
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && isb % 8 == 0 && ((uintptr_t)in & 15) == 0,
               "lowbit_quant_per_thread: input base address and strides must keep 16-byte alignment");
  LOWBIT_CHECK(osn % 8 == 0 && osh % 8 == 0 && osb % 8 == 0, "lowbit_quant_per_thread: output strides misaligned");
  const int ng = is_key ? 4 : 8;
  LOWBIT_CHECK(n_scale % ng == 0 && n_scale / ng >= (N + warp_blk - 1) / warp_blk, "lowbit_quant_per_thread: n_scale too small");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n_scale / ng, H, B);
#define LAUNCH(T, DD)                                                                                                  \
  if (is_key) quant_per_thread_kernel<T, DD, 64, true><<<grid, 256, 0, st>>>((const T*)in, (const T*)km, (int8_t*)codes, scale, N, n_scale, isb, ish, isn, osb, osh, osn, bits, H); \
  else quant_per_thread_kernel<T, DD, 32, false><<<grid, 256, 0, st>>>((const T*)in, (const T*)km, (int8_t*)codes, scale, N, n_scale, isb, ish, isn, osb, osh, osn, bits, H);
  if (dtype == LOWBIT_F16) {
    if (D == 64) { LAUNCH(__half, 64) } else { LAUNCH(__half, 128) }
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) { LAUNCH(__nv_bfloat16, 64) } else { LAUNCH(__nv_bfloat16, 128) }
  } else {
    return fail("lowbit_quant_per_thread: unsupported dtype %d", dtype);
  }
#undef LAUNCH
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int lowbit_quant_pack_lastdim(const void* data, void* code, void* scale, void* mn, int64_t rows, int T, int group,
                              int bits, int dtype, void* stream) {
  LOWBIT_CHECK(data && code && scale && mn, "lowbit_quant_pack_lastdim: null pointer");
  LOWBIT_CHECK(group == 32, "lowbit_quant_pack_lastdim: group_size must be 32 (got %d)", group);
  LOWBIT_CHECK(bits == 2 || bits == 4 || bits == 8, "lowbit_quant_pack_lastdim: bits must be 2, 4 or 8 (got %d)", bits);
  LOWBIT_CHECK(T > 0 && T % group == 0, "lowbit_quant_pack_lastdim: T must be a positive multiple of group_size");
  LOWBIT_CHECK(dtype == LOWBIT_F16, "lowbit_quant_pack_lastdim: only float16 data is supported");
  LOWBIT_CHECK(((uintptr_t)data & 15) == 0 && ((uintptr_t)code & 3) == 0, "lowbit_quant_pack_lastdim: misaligned pointer");
  if (rows == 0) return 0;
  const int64_t ng = rows * (T / group);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((ng + 255) / 256);
  if (bits == 2) kivi_pack_kernel<2><<<blocks, 256, 0, st>>>((const __half*)data, (uint8_t*)code, (__half*)scale, (__half*)mn, ng);
  else if (bits == 4) kivi_pack_kernel<4><<<blocks, 256, 0, st>>>((const __half*)data, (uint8_t*)code, (__half*)scale, (__half*)mn, ng);
  else kivi_pack_kernel<8><<<blocks, 256, 0, st>>>((const __half*)data, (uint8_t*)code, (__half*)scale, (__half*)mn, ng);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

int64_t lowbit_v_fp8_workspace_bytes(int B, int H, int N, int D) {
  (void)N;
  return (int64_t)B * H * kVChunks * D * (int64_t)sizeof(VPartial) + (int64_t)B * H * D * 8;  // chunk statistics + (1/scale, mean) table
}

int lowbit_v_fp8_per_channel(const void* v, void* v8, float* v_scale, float* vm, void* workspace, int B, int H, int N,
                             int D, int64_t sb, int64_t sh, int64_t sn, int64_t osb, int64_t osh, int64_t osd,
                             float scale_max, int dtype, void* stream) {
  LOWBIT_CHECK(v && v8 && v_scale && workspace, "lowbit_v_fp8_per_channel: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_v_fp8_per_channel: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_v_fp8_per_channel: empty tensor");
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0 && ((uintptr_t)v & 15) == 0,
               "lowbit_v_fp8_per_channel: input base address and strides must keep 16-byte alignment");
  LOWBIT_CHECK(osb % 16 == 0 && osh % 16 == 0 && osd % 16 == 0 && ((uintptr_t)v8 & 15) == 0,
               "lowbit_v_fp8_per_channel: output must keep 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  int chunk = (N + kVChunks - 1) / kVChunks;
  chunk = (chunk + 63) / 64 * 64;
  const int nchunk = (N + chunk - 1) / chunk;
  const int npad = (N + 63) / 64 * 64;
  dim3 g1(nchunk, H, B);
  const int total = B * H * D;
  float2* tab = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(workspace) + (int64_t)B * H * nchunk * D * sizeof(VPartial));
#define LAUNCH(T, DD)                                                                                               \
  if (vm != nullptr) v_stats_partial_kernel<T, DD, true><<<g1, 256, 0, st>>>((const T*)v, (VPartial*)workspace, N, chunk, nchunk, sb, sh, sn, H); \
  else v_stats_partial_kernel<T, DD, false><<<g1, 256, 0, st>>>((const T*)v, (VPartial*)workspace, N, chunk, nchunk, sb, sh, sn, H); \
  v_stats_final_kernel<T, DD><<<(total + 255) / 256, 256, 0, st>>>((const VPartial*)workspace, tab, v_scale, vm, N, nchunk, scale_max, total); \
  v_fp8_quant_kernel<T, DD><<<dim3((N + kVTiles * VQuantCfg<DD>::TT - 1) / (kVTiles * VQuantCfg<DD>::TT), H, B), 256, 0, st>>>( \
      (const T*)v, tab, (uint8_t*)v8, N, npad, sb, sh, sn, osb, osh, osd, H);
  if (dtype == LOWBIT_F16) {
    if (D == 64) { LAUNCH(__half, 64) } else { LAUNCH(__half, 128) }
  } else if (dtype == LOWBIT_BF16) {
    if (D == 64) { LAUNCH(__nv_bfloat16, 64) } else { LAUNCH(__nv_bfloat16, 128) }
  } else {
    return fail("lowbit_v_fp8_per_channel: unsupported dtype %d", dtype);
  }
#undef LAUNCH
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}


int lowbit_sub_mean(const void* v, const void* vm, void* out, int B, int H, int N, int D, int64_t sb, int64_t sh,
                    int64_t sn, int64_t osb, int64_t osh, int64_t osn, int dtype, void* stream) {
  LOWBIT_CHECK(v && vm && out, "lowbit_sub_mean: null pointer");
  LOWBIT_CHECK(D > 0 && D % 8 == 0, "lowbit_sub_mean: head_dim must be a multiple of 8 (got %d)", D);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_sub_mean: empty tensor");
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0 && osn % 8 == 0 && osh % 8 == 0 && osb % 8 == 0,
               "lowbit_sub_mean: strides must keep 16-byte alignment");
  LOWBIT_CHECK(((uintptr_t)v & 15) == 0 && ((uintptr_t)vm & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "lowbit_sub_mean: pointers must be 16-byte aligned");
  const int64_t total8 = (int64_t)B * H * N * (D / 8);
  const unsigned blocks = (unsigned)((total8 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LOWBIT_F16)
    sub_mean_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)v, (const __half*)vm, (__half*)out, H, N, D, sb, sh, sn, osb, osh, osn, total8);
  else if (dtype == LOWBIT_BF16)
    sub_mean_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)v, (const __nv_bfloat16*)vm, (__half*)out, H, N, D, sb, sh, sn, osb, osh, osn, total8);
  else
    return fail("lowbit_sub_mean: unsupported dtype %d", dtype);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
