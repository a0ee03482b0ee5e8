// common.cuh -- shared host/device helpers for liblowbit_fa_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lowbit_fa.h"

namespace lowbit {

// thread-local error string behind lowbit_last_error()
std::string& last_error();
int fail(const char* fmt, ...);

#define LOWBIT_CHECK(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::lowbit::fail(__VA_ARGS__);              \
  } while (0)

#define LOWBIT_CUDA(call)                                                                        \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return ::lowbit::fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// host: 4-D TMA tensor map over a strided [d3][d2][d1][d0] view (defined in attn.cu): dims (d0 innermost .. d3),
// element strides of dims 1..3, box (box0, box1, 1, 1); returns non-zero (and sets last_error) on failure
int make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, const int64_t (&dim)[4],
             const int64_t (&stride_elems)[3], int box0, int box1, CUtensorMapSwizzle swz);

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 128-bit streaming global load (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// four predicated 128-bit streaming loads in one asm statement (a predicate that is off leaves zeros).  NB: this alone
// does not keep them in flight together -- ptxas still sinks each load to its first use and recycles one destination quad
// (load -> use -> load) when it is short of registers; the callers' __launch_bounds__(256, 2) is what gives it the room
__device__ __forceinline__ void ld_stream_v4_x4(const void* p0, const void* p1, const void* p2, const void* p3,
                                                bool k0, bool k1, bool k2, bool k3, uint4 (&r)[4]) {
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2, q3;\n\t"
      "setp.ne.u32 q0, %20, 0;\n\tsetp.ne.u32 q1, %21, 0;\n\tsetp.ne.u32 q2, %22, 0;\n\tsetp.ne.u32 q3, %23, 0;\n\t"
      "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0;\n\t"
      "mov.u32 %4, 0; mov.u32 %5, 0; mov.u32 %6, 0; mov.u32 %7, 0;\n\t"
      "mov.u32 %8, 0; mov.u32 %9, 0; mov.u32 %10, 0; mov.u32 %11, 0;\n\t"
      "mov.u32 %12, 0; mov.u32 %13, 0; mov.u32 %14, 0; mov.u32 %15, 0;\n\t"
      "@q0 ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%16];\n\t"
      "@q1 ld.global.nc.L1::no_allocate.v4.u32 {%4,%5,%6,%7}, [%17];\n\t"
      "@q2 ld.global.nc.L1::no_allocate.v4.u32 {%8,%9,%10,%11}, [%18];\n\t"
      "@q3 ld.global.nc.L1::no_allocate.v4.u32 {%12,%13,%14,%15}, [%19];\n\t}"
      : "=&r"(r[0].x), "=&r"(r[0].y), "=&r"(r[0].z), "=&r"(r[0].w), "=&r"(r[1].x), "=&r"(r[1].y), "=&r"(r[1].z), "=&r"(r[1].w),
        "=&r"(r[2].x), "=&r"(r[2].y), "=&r"(r[2].z), "=&r"(r[2].w), "=&r"(r[3].x), "=&r"(r[3].y), "=&r"(r[3].z), "=&r"(r[3].w)
      : "l"(p0), "l"(p1), "l"(p2), "l"(p3), "r"((uint32_t)k0), "r"((uint32_t)k1), "r"((uint32_t)k2), "r"((uint32_t)k3));
}

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = to_f32<T>(h[i]);
}

// fp32 pair -> the input dtype (one packed conversion, round to nearest even) -> fp32 pair: `k - km` as the framework
// computes it in the tensor's dtype (quant_per_block.py:186-187)
template <typename T> __device__ __forceinline__ float2 round_trip2(float2 v);
template <> __device__ __forceinline__ float2 round_trip2<__half>(float2 v) { return __half22float2(__float22half2_rn(v)); }
template <> __device__ __forceinline__ float2 round_trip2<__nv_bfloat16>(float2 v) {
  return __bfloat1622float2(__float22bfloat162_rn(v));
}

// Load side of the per-block quantizers for one 8-element row chunk, on packed fp32 pairs (FADD2 / FMUL2: one issue
// slot per two elements -- these kernels are instruction-issue bound, not HBM bound, at 15-30 us per launch):
//   x = f32(in) [- f32(km)] [rounded to the input dtype] * sm,  amax = max(amax, |x|)
// A row that is out of range must contribute exact zeros: without km that is what the zero-filled load gives
// (0 * sm); with km the caller substitutes km's own bits for the row (km - km = 0), which is cheaper than a select
// per element.  Same IEEE operations per lane as the scalar form, so the same bits.
// packed pair types of the input dtypes, and `a - b` in that dtype (one correctly rounded HSUB2 per pair)
template <typename T> struct Pair;
template <> struct Pair<__half> {
  using type = __half2;
  static __device__ __forceinline__ type from_f32(float2 v) { return __float22half2_rn(v); }
  static __device__ __forceinline__ float2 to_f32(type v) { return __half22float2(v); }
};
template <> struct Pair<__nv_bfloat16> {
  using type = __nv_bfloat162;
  static __device__ __forceinline__ type from_f32(float2 v) { return __float22bfloat162_rn(v); }
  static __device__ __forceinline__ float2 to_f32(type v) { return __bfloat1622float2(v); }
};

template <typename T, bool HAS_KM, bool ROUND_KM>
__device__ __forceinline__ void prep_row8(const uint4& raw, const float (&kmf)[8], float sm, float (&x)[8], float& amax) {
  const float2 sm2 = make_float2(sm, sm);
  if constexpr (HAS_KM && ROUND_KM) {
    // `k - km` in the tensor's dtype is ONE subtraction there: RN_dtype(a - b).  The fp32 route below (convert,
    // subtract, round back) gives the same bits -- fp32's 24 bits are >= 2 * 11 + 2, so the double rounding is
    // innocuous -- at twice the instructions; km comes back to its packed form exactly (hoisted out of the row loop).
    using P = Pair<T>;
    const typename P::type* h2 = reinterpret_cast<const typename P::type*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const typename P::type km2 = P::from_f32(make_float2(kmf[2 * i], kmf[2 * i + 1]));
      float2 v = P::to_f32(__hsub2(h2[i], km2));
      v = __fmul2_rn(v, sm2);
      x[2 * i] = v.x;
      x[2 * i + 1] = v.y;
      amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
    }
    return;
  }
  float f[8];
  unpack8<T>(raw, f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 v = make_float2(f[2 * i], f[2 * i + 1]);
    if constexpr (HAS_KM) {
      v = __fadd2_rn(v, make_float2(-kmf[2 * i], -kmf[2 * i + 1]));  // x - km == x + (-km), bit for bit
      if constexpr (ROUND_KM) v = round_trip2<T>(v);
    }
    v = __fmul2_rn(v, sm2);
    x[2 * i] = v.x;
    x[2 * i + 1] = v.y;
    amax = fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y)));
  }
}

// Exact per-thread accumulation.  fp16: every value is an integer X = x * 2^24 with |X| < 2^40; it is split without any
// conversion instruction into hi = RN(16 x) (an integer-valued float, |hi| < 2^20) and lo = x - hi / 16 (a multiple of
// 2^-24, |lo| <= 2^-5) by magic-constant arithmetic (1.5 * 2^23) on packed fp32 pairs, and both parts are summed IN
// FP32 -- exactly, for up to 16 rows (|sum hi| < 2^24, |sum lo * 2^24| < 2^24) -- then folded into two int32 lanes
// (two F2I per column and 16 rows: the caller calls fold() at least every 16 add()s), which flush() moves to int64
// at least every 256 rows.  3.5 instructions per element instead of the 7 of the first version (raw float bits
// summed in integer lanes: two IADD per element).  bf16: fp64 accumulation in a fixed order.
template <typename T> struct RowAcc;
template <> struct RowAcc<__half> {
  float2 hf[4], lf[4];
  int hi[8], lo[8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) hf[i] = lf[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) hi[i] = lo[i] = 0;
  }
  __device__ __forceinline__ void fold() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hi[2 * i] += __float2int_rn(hf[i].x);
      hi[2 * i + 1] += __float2int_rn(hf[i].y);
      lo[2 * i] += __float2int_rn(lf[i].x * 16777216.0f);
      lo[2 * i + 1] += __float2int_rn(lf[i].y * 16777216.0f);
      hf[i] = lf[i] = make_float2(0.f, 0.f);
    }
  }
  __device__ __forceinline__ void add(const uint4& raw) {
    const __half2* hv = reinterpret_cast<const __half2*>(&raw);
    const float kM = 12582912.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __half22float2(hv[i]);
      const float2 a = __ffma2_rn(x, make_float2(16.0f, 16.0f), make_float2(kM, kM));
      const float2 t = __fadd2_rn(a, make_float2(-kM, -kM));                   // RN(16 x), exact
      const float2 rem = __ffma2_rn(t, make_float2(-0.0625f, -0.0625f), x);   // exact: |rem| <= 2^-5, multiple of 2^-24
      hf[i] = __fadd2_rn(hf[i], t);
      lf[i] = __fadd2_rn(lf[i], rem);
    }
  }
  __device__ __forceinline__ void flush(long long (&acc)[8]) {
    fold();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += (long long)hi[i] * 1048576ll + (long long)lo[i];
      hi[i] = lo[i] = 0;
    }
  }
};
template <> struct RowAcc<__nv_bfloat16> {
  double s[8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.0;
  }
  __device__ __forceinline__ void add(const uint4& raw) {
    const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += (double)__bfloat162float(hv[i]);
  }
  __device__ __forceinline__ void fold() {}
  __device__ __forceinline__ void flush(double (&acc)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += s[i];
    clear();
  }
};

__device__ __forceinline__ float warp_max(float v) {
  // all operands are >= 0, so unsigned integer order == float order
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

}  // namespace lowbit
