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
template <int NPASS>
__device__ __forceinline__ void quantize_rows(const float (&x)[NPASS][8], int8_t* dst, int64_t osn, int row0, int rpp,
                                              int r0, int c8, int blk, int N, float sc, float rcp, bool triton,
                                              bool slow_div, int bits, int pack, bool gpu_div = false) {
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const int rl = p * rpp + r0;
    const int row = row0 + rl;
    if (!(rl < blk && row < N)) continue;
    int c[8];
    if (triton && gpu_div) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = q1_code_divfull(x[p][i], sc);
    } else if (triton && !slow_div) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = q1_code_fast(x[p][i], sc, rcp);
    } else if (triton) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = q1_code_ieee(x[p][i], sc);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) c[i] = max(-128, min(127, __float2int_rn(__fmul_rn(x[p][i], rcp))));
    }
    store_codes8(dst + (int64_t)row * osn, c8, c, bits, pack);
  }
}

template <typename T, int D, int BLK>
__global__ void __launch_bounds__(kQuantThreads)
quant_per_block_kernel(const T* __restrict__ in, const T* __restrict__ km, int8_t* __restrict__ out,
                       float* __restrict__ scale, int N, int nblk,
                       int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                       float sm, int bits, int pack, int mode, int H) {
  constexpr int TPR = D / 8;                 // threads per row
  constexpr int RPP = kQuantThreads / TPR;   // rows per pass
  constexpr int NP = (BLK + RPP - 1) / RPP;  // passes
  const int tid = threadIdx.x;
  const int c8 = (tid % TPR) * 8;
  const int r0 = tid / TPR;
  const int jb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const T* src = in + b * isb + h * ish + c8;

  float kmf[8];
  const bool has_km = km != nullptr;
  if (has_km) {
    uint4 raw = *reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8);
    unpack8<T>(raw, kmf);
  }

  float x[NP][8];
  uint4 raw[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const int rl = p * RPP + r0;
    const int row = jb * BLK + rl;
    raw[p] = make_uint4(0, 0, 0, 0);
    if (rl < BLK && row < N) raw[p] = ld_stream_v4(src + (int64_t)row * isn);
  }
  float amax = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    unpack8<T>(raw[p], x[p]);
    const int rl = p * RPP + r0;
    const bool live = (rl < BLK) && (jb * BLK + rl < N);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = x[p][i];
      if (has_km) {
        v = __fsub_rn(v, kmf[i]);
        if ((mode & 0xff) == LOWBIT_QMODE_TRITON) v = to_f32<T>(from_f32<T>(v));  // `k - km` in the input dtype
      }
      v = __fmul_rn(v, sm);
      v = live ? v : 0.f;  // rows >= N contribute 0 (masked load), even when km != 0
      x[p][i] = v;
      amax = fmaxf(amax, fabsf(v));
    }
  }
  // block abs-max: warp redux + smem
  __shared__ float s_w[kQuantThreads / 32];
  amax = warp_max(amax);
  if ((tid & 31) == 0) s_w[tid >> 5] = amax;
  __syncthreads();
  float bmax = s_w[0];
#pragma unroll
  for (int w = 1; w < kQuantThreads / 32; ++w) bmax = fmaxf(bmax, s_w[w]);

  const float qmax = bits == 8 ? 127.f : (bits == 4 ? 7.f : 1.f);
  float sc, rcp = 0.f;
  const bool triton = (mode & 0xff) == LOWBIT_QMODE_TRITON;
  bool slow_div = (mode & LOWBIT_QMODE_FLAG_IEEE_DIV) != 0;
  const bool gpu_div = triton && (mode & LOWBIT_QMODE_FLAG_DIV_FULL) != 0;
  if (gpu_div) {
    sc = div_full(bmax, qmax);
  } else if (triton) {
    sc = __fdiv_rn(bmax, qmax);
    rcp = __frcp_rn(sc);  // correctly rounded 1/scale, once per block
    // bf16 blocks near the ends of the fp32 exponent range (1/scale denormal or infinite), and all-zero blocks,
    // take the IEEE division: a block-uniform branch
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

template <typename T, int D, int BLK>
static int launch_qpb(const void* in, const void* km, void* codes, float* scale, int B, int H, int N,
                      int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                      float sm, int bits, int pack, int mode, cudaStream_t st) {
  const int nblk = (N + BLK - 1) / BLK;
  dim3 grid(nblk, H, B);
  quant_per_block_kernel<T, D, BLK><<<grid, kQuantThreads, 0, st>>>(
      (const T*)in, (const T*)km, (int8_t*)codes, scale, N, nblk, isb, ish, isn, osb, osh, osn, sm, bits, pack, mode, H);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
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

template <typename T, int D, int BLK>
__global__ void __launch_bounds__(kQuantThreads)
quant_per_block_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const T* __restrict__ km,
                           int8_t* __restrict__ out, float* __restrict__ scale, int N, int nblk, int H, int total,
                           int64_t osb, int64_t osh, int64_t osn, float sm, int bits, int pack, int mode) {
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
  auto issue = [&](int t, int stage) {  // thread 0: fetch tile t (linear over b, h, block) into `stage`
    const int j = t % nblk, h = (t / nblk) % H, b = t / (nblk * H);
    ptx::mbar_expect_tx(full + stage, C::kTileBytes);
    ptx::tma_load_4d(smem + stage * C::kTileBytes, &tmIn, full + stage, 0, j * BLK, h, b);
  };
  const int first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
    for (int i = 0; i < S; ++i)
      if (first + i * stride < total) issue(first + i * stride, i);
  }
  int it = 0;
  for (int t = first; t < total; t += stride, ++it) {
    const int stage = it % S;
    const int jb = t % nblk, h = (t / nblk) % H, b = t / (nblk * H);
    float kmf[8];
    const bool has_km = km != nullptr;
    if (has_km) {
      uint4 raw = *reinterpret_cast<const uint4*>(km + ((int64_t)b * H + h) * D + c8);
      unpack8<T>(raw, kmf);
    }
    ptx::mbar_wait(full + stage, (it / S) & 1, 40);
    const uint8_t* tile = smem + stage * C::kTileBytes;
    float x[NP][8];
    float amax = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int rl = p * RPP + r0;
      const bool live = (rl < BLK) && (jb * BLK + rl < N);
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (rl < BLK) raw = *reinterpret_cast<const uint4*>(tile + ((size_t)rl * D + c8) * 2);
      unpack8<T>(raw, x[p]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = x[p][i];
        if (has_km) {
          v = __fsub_rn(v, kmf[i]);
          if (triton) v = to_f32<T>(from_f32<T>(v));  // `k - km` in the input dtype
        }
        v = __fmul_rn(v, sm);
        v = live ? v : 0.f;  // rows >= N contribute 0 (masked load), even when km != 0
        x[p][i] = v;
        amax = fmaxf(amax, fabsf(v));
      }
    }
    amax = warp_max(amax);
    if ((tid & 31) == 0) s_w[it & 1][tid >> 5] = amax;
    __syncthreads();  // every thread has finished reading this stage
    if (tid == 0 && t + S * stride < total) issue(t + S * stride, stage);
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

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
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
  auto kern = quant_per_block_tma_kernel<T, D, BLK>;
  static bool configured = false;
  if (!configured) {
    LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
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

template <typename T, int D>
__global__ void __launch_bounds__(256)
k_mean_partial_kernel(const T* __restrict__ k, typename MeanAcc<T>::type* __restrict__ part, int N, int chunk,
                      int nchunk, int64_t sb, int64_t sh, int64_t sn, int H) {
  using A = typename MeanAcc<T>::type;
  constexpr int TPR = D / 8, RPP = 256 / TPR;
  const int tid = threadIdx.x, c8 = (tid % TPR) * 8, r0 = tid / TPR;
  const int ch = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
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
    uint4 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      raw[u] = make_uint4(0, 0, 0, 0);
      if (row + u * RPP < row_end) raw[u] = ld_stream_v4(src + (int64_t)(row + u * RPP) * sn);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) ra.add(raw[u]);
    since_flush += 8;
    if (since_flush >= 256) { ra.flush(acc); since_flush = 0; }
  }
  ra.flush(acc);
  __shared__ A s[RPP][D + 1];
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

}  // namespace lowbit

using namespace lowbit;

extern "C" {

int lowbit_version(void) { return LOWBIT_ABI_VERSION; }
const char* lowbit_last_error(void) { return last_error().c_str(); }

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
  LOWBIT_CHECK(sn % 8 == 0 && sh % 8 == 0 && sb % 8 == 0, "lowbit_k_mean: strides must keep 16-byte alignment");
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

int lowbit_quant_per_block(const void* in, const void* km, void* codes, float* scale, int B, int H, int N, int D,
                           int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                           int blk, int bits, int pack, float sm, int mode, int dtype, void* stream) {
  LOWBIT_CHECK(in && codes && scale, "lowbit_quant_per_block: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_quant_per_block: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(blk == 32 || blk == 64 || blk == 128, "lowbit_quant_per_block: blk must be 32, 64 or 128 (got %d)", blk);
  LOWBIT_CHECK(bits == 8 || bits == 4 || bits == 2, "lowbit_quant_per_block: bits must be 8, 4 or 2 (got %d)", bits);
  LOWBIT_CHECK((mode & 0xff) == LOWBIT_QMODE_TRITON || (mode & 0xff) == LOWBIT_QMODE_CUDA, "lowbit_quant_per_block: bad mode %d", mode);
  LOWBIT_CHECK(B > 0 && H > 0 && N > 0, "lowbit_quant_per_block: empty tensor");
  LOWBIT_CHECK(isn % 8 == 0 && ish % 8 == 0 && isb % 8 == 0, "lowbit_quant_per_block: input strides must keep 16-byte alignment");
  const int ob = (pack && bits < 8) ? bits : 8;  // bytes of 8 codes
  LOWBIT_CHECK(osn % ob == 0 && osh % ob == 0 && osb % ob == 0, "lowbit_quant_per_block: output strides misaligned");
  cudaStream_t st = (cudaStream_t)stream;
#define ARGS in, km, codes, scale, B, H, N, isb, ish, isn, osb, osh, osn, sm, bits, pack, mode, st
  // 128-row blocks go through the TMA-pipelined persistent kernel, 64-row blocks through the one-block-per-CTA
  // kernel (measured faster at that tile size: 8-16 KB tiles do not amortise the per-tile barrier).
  // LOWBIT_QUANT_TMA=0 / 2 force neither / both (A/B measurements).
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("LOWBIT_QUANT_TMA"); use_tma = e ? atoi(e) : 1; }
  const bool tma_ok = use_tma && ((uintptr_t)in & 15) == 0;
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
