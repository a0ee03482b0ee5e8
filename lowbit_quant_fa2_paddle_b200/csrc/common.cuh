// common.cuh -- shared host/device helpers for liblowbit_fa_b200.so (sm_100a only).
#pragma once
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

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = to_f32<T>(h[i]);
}

__device__ __forceinline__ float warp_max(float v) {
  // all operands are >= 0, so unsigned integer order == float order
  return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

}  // namespace lowbit
