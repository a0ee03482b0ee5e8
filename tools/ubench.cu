// ubench.cu -- instruction-throughput microbenchmarks behind the softmax design choices in csrc/attn.cu
// (development aid, not part of the library).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
// Prints thread-ops per clock per SM for each instruction class, measured with %clock64 on a full grid.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define OPS 32
template <int KIND>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int iters, float seed) {
  float f[OPS];
  uint32_t u[OPS];
#pragma unroll
  for (int i = 0; i < OPS; ++i) { f[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; u[i] = __float_as_uint(f[i]); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < OPS; ++i) {
      if (KIND == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (KIND == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (KIND == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (KIND == 3) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(u[i])); u[i] = __float_as_uint(f[i]); }
      if (KIND == 4) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) % OPS])); f[i] = __uint_as_float(u[i]); }
      if (KIND == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(seed), "f"(f[(i + 1) % OPS]));
      if (KIND == 6) { if (i % 2 == 0) { unsigned long long a, b, c;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(f[i + 1]));
          asm volatile("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(seed));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a) : "l"(b));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(f[i + 1]) : "l"(a)); (void)c; } }
      if (KIND == 7) asm volatile("max.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % OPS]));
      if (KIND == 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % OPS]));
      if (KIND == 9) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % OPS]), "r"(u[(i + 2) % OPS]));
      if (KIND == 10) asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(*(uint16_t*)&u[i]) : "f"(f[i]), "f"(f[(i + 1) % OPS]));
      if (KIND == 11) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % OPS]));
      if (KIND == 12) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*(uint16_t*)&u[i]));
      if (KIND == 13) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % OPS]), "r"(u[(i + 2) % OPS]));
      if (KIND == 14) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % OPS]));
      if (KIND == 15) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % OPS]));
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < OPS; ++i) acc += f[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the softmax inner block of csrc/attn.cu without TMEM: 32 scores -> 16 packed fp16 words + row sum
template <int MODE>
__global__ void __launch_bounds__(128) mix(float* out, long long* cyc, int iters, float sc, int* in) {
  int s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[i] + threadIdx.x;
  float l = 0.f;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
    int mx = -2147483647;
    if (MODE & 1) {
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = max(mx, s[c]);
    }
    const float nm = -(float)mx * sc * 1e-9f;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      float2 x0 = __ffma2_rn(make_float2(__int2float_rn(s[c]), __int2float_rn(s[c + 1])), make_float2(sc, sc), make_float2(nm, nm));
      float2 x1 = __ffma2_rn(make_float2(__int2float_rn(s[c + 2]), __int2float_rn(s[c + 3])), make_float2(sc, sc), make_float2(nm, nm));
      float2 p0, p1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0.x) : "f"(x0.x));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0.y) : "f"(x0.y));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1.x) : "f"(x1.x));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1.y) : "f"(x1.y));
      if (MODE & 2) { a0 = __fadd2_rn(a0, p0); a1 = __fadd2_rn(a1, p1); }
      uint32_t w0, w1;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(p0.y), "f"(p0.x));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(p1.y), "f"(p1.x));
      acc ^= w0 ^ w1;
    }
    l += a0.x + a0.y + a1.x + a1.y;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] += (int)(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}



// Row-sum / int->float alternatives for the same block (development A/B):
//  SUM: 0 none, 1 FADD2 x2 accumulators (current), 2 FADD2 x4 accumulators, 3 scalar FADD x4, 4 HADD2 on the packed words (fp16 accumulate),
//       5 HADD2 pairs then HADD2.F32-style widening every 4 words
//  CVT: 0 I2FP (current), 1 integer magic add folded into the FFMA constant (IADD + FFMA2, no conversion pipe)
template <int SUM, int CVT>
__global__ void __launch_bounds__(128) mix2(float* out, long long* cyc, int iters, float sc, int* in) {
  int s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[i] + threadIdx.x;
  float l = 0.f;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
    uint32_t h0 = 0, h1 = 0;
    int mx = -2147483647;
#pragma unroll
    for (int c = 0; c < 32; ++c) mx = max(mx, s[c]);
    float nm = -(float)mx * sc * 1e-9f;
    if (CVT == 1) nm = nm - 12582912.0f * sc;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      float v0, v1, v2, v3;
      if (CVT == 0) { v0 = __int2float_rn(s[c]); v1 = __int2float_rn(s[c + 1]); v2 = __int2float_rn(s[c + 2]); v3 = __int2float_rn(s[c + 3]); }
      else { v0 = __int_as_float(s[c] + 0x4B400000); v1 = __int_as_float(s[c + 1] + 0x4B400000);
             v2 = __int_as_float(s[c + 2] + 0x4B400000); v3 = __int_as_float(s[c + 3] + 0x4B400000); }
      float2 x0 = __ffma2_rn(make_float2(v0, v1), make_float2(sc, sc), make_float2(nm, nm));
      float2 x1 = __ffma2_rn(make_float2(v2, v3), make_float2(sc, sc), make_float2(nm, nm));
      float2 p0, p1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0.x) : "f"(x0.x));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0.y) : "f"(x0.y));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1.x) : "f"(x1.x));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1.y) : "f"(x1.y));
      if (SUM == 1) { a0 = __fadd2_rn(a0, p0); a1 = __fadd2_rn(a1, p1); }
      if (SUM == 2) { if (c % 8 == 0) { a0 = __fadd2_rn(a0, p0); a1 = __fadd2_rn(a1, p1); } else { a2 = __fadd2_rn(a2, p0); a3 = __fadd2_rn(a3, p1); } }
      if (SUM == 3) { f0 += p0.x; f1 += p0.y; f2 += p1.x; f3 += p1.y; }
      uint32_t w0, w1;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(p0.y), "f"(p0.x));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(p1.y), "f"(p1.x));
      if (SUM == 4 || SUM == 5) {
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h0) : "r"(w0));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h1) : "r"(w1));
        if (SUM == 5 && (c % 16 == 12)) {  // widen every 8 words
          uint32_t hs; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(hs) : "r"(h0), "r"(h1));
          float lo, hi;
          asm volatile("{.reg .b16 a, b; mov.b32 {a, b}, %2; cvt.f32.f16 %0, a; cvt.f32.f16 %1, b;}" : "=f"(lo), "=f"(hi) : "r"(hs));
          f0 += lo; f1 += hi; h0 = 0; h1 = 0;
        }
      }
      acc ^= w0 ^ w1;
    }
    if (SUM == 4) { uint32_t hs; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(hs) : "r"(h0), "r"(h1));
      float lo, hi; asm volatile("{.reg .b16 a, b; mov.b32 {a, b}, %2; cvt.f32.f16 %0, a; cvt.f32.f16 %1, b;}" : "=f"(lo), "=f"(hi) : "r"(hs)); f0 += lo + hi; }
    l += a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + f0 + f1 + f2 + f3;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] += (int)(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// phase-structured variant: (A) convert+scale all 32 in place, (B) 32 MUFU back to back, (C) pack + row sum.
// MODE bit0: integer row max first; bit1: fp32 row sum; bit2: no max but a (never taken) check of the block sum.
template <int MODE>
__global__ void __launch_bounds__(128) mixp(float* out, long long* cyc, int iters, float sc, int* in) {
  int s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[i] + threadIdx.x;
  float l = 0.f;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    int mx = -2147483647;
    if (MODE & 1) {
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = max(mx, s[c]);
    }
    const float nm = -(float)mx * sc * 1e-9f;
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float2 t = __ffma2_rn(make_float2(__int2float_rn(s[c]), __int2float_rn(s[c + 1])), make_float2(sc, sc), make_float2(nm, nm));
      x[c] = t.x; x[c + 1] = t.y;
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
    float2 a0 = make_float2(0.f, 0.f), a1 = a0;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      if (MODE & 2) { a0 = __fadd2_rn(a0, make_float2(x[c], x[c + 1])); a1 = __fadd2_rn(a1, make_float2(x[c + 2], x[c + 3])); }
      uint32_t w0, w1;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(x[c + 1]), "f"(x[c]));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(x[c + 3]), "f"(x[c + 2]));
      acc ^= w0 ^ w1;
    }
    const float lb = a0.x + a0.y + a1.x + a1.y;
    if (MODE & 4) { if (__any_sync(0xffffffffu, lb > 1e30f)) acc ^= 0x55; }
    l += lb;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] += (int)(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// pairwise interference: 32 MUFU.EX2 + 32 ops of another class per iteration, independent chains
template <int KIND>
__global__ void __launch_bounds__(128) combo(float* out, long long* cyc, int iters, float seed) {
  float f[32], g[32];
  uint32_t u[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { f[i] = seed + i * 0.001f; g[i] = seed * i + threadIdx.x; u[i] = __float_as_uint(g[i]); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (KIND == 0) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(g[i]) : "r"(u[i])); u[i] = __float_as_uint(g[i]); }
      if (KIND == 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(g[i]), "f"(g[(i + 1) % 32])); g[i] = __uint_as_float(u[i]); }
      if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(seed), "f"(g[(i + 1) % 32]));
      if (KIND == 3) asm volatile("max.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % 32]));
      if (KIND == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(g[(i + 1) % 32]));
      if (KIND == 5) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(g[i]) : "r"(u[i])); u[i] = __float_as_uint(g[i]);
                       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(seed), "f"(g[(i + 1) % 32])); }
      if (KIND == 6) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % 32]), "r"(u[(i + 2) % 32])); }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += f[i] + g[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int blocks_per_sm, double ops_per_thread_iter, int iters) {
  int nsm = 148;
  int grid = nsm * blocks_per_sm;
  float* out; long long* cyc;
  cudaMalloc(&out, grid * 128 * sizeof(float));
  cudaMalloc(&cyc, grid * sizeof(long long));
  launch(grid, out, cyc, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(grid, out, cyc, iters);
  cudaEventRecord(e0);
  launch(grid, out, cyc, iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid];
  cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  double per_sm = blocks_per_sm * 128.0 * ops_per_thread_iter * iters / avg;
  double wall = (double)grid * 128.0 * ops_per_thread_iter * iters / (ms * 1e-3) / 148.0 / 1.965e9;
  printf("%-28s warps/SM %2d  %8.2f thread-ops/tick/SM  %7.2f thread-ops/clk/SM @1965MHz (%.3f ms, %s)\n", name, blocks_per_sm * 4, per_sm, wall, ms, cudaGetErrorString(e));
  delete[] h; cudaFree(out); cudaFree(cyc);
}

int main() {
  const char* names[] = {"ex2.f32", "ex2.f16x2 (x2 elems)", "ex2.bf16x2 (x2 elems)", "cvt.f32.s32 (I2FP)", "cvt.f16x2.f32 (F2FP)",
                         "fma.f32", "fma.f32x2 (per pair)", "max.s32", "add.s32", "mad.lo.s32", "cvt.e4m3x2.f32",
                         "add.f16x2", "ex2.f16", "lop3", "max.f32", "add.f32"};
  for (int bps : {8, 16}) {
#define R(K, OPI) run(names[K], [](int g, float* o, long long* c, int it) { k<K><<<g, 128>>>(o, c, it, 0.5f); }, bps, OPI, 20000);
    R(0, 32) R(3, 32) R(4, 32) R(5, 32) R(6, 16) R(7, 32) R(8, 32) R(11, 32) R(13, 32) R(15, 32)
  }
  {
    const char* cn[] = {"ex2 + I2FP", "ex2 + F2FP", "ex2 + FFMA", "ex2 + IMNMX", "ex2 + FADD", "ex2 + I2FP + FFMA", "ex2 + IMAD"};
#define C(K) run(cn[K], [](int g, float* o, long long* c, int it) { combo<K><<<g, 128>>>(o, c, it, 0.5f); }, 8, 32, 20000);
    C(0) C(1) C(2) C(3) C(4) C(5) C(6)
  }
  int* in; cudaMalloc(&in, 32 * sizeof(int)); cudaMemset(in, 1, 32 * sizeof(int));
  for (int bps : {4, 8}) {
    run("mix: cvt+ffma2+ex2+pack", [=](int g, float* o, long long* c, int it) { mix<0><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    run("mix + fadd2 row sum", [=](int g, float* o, long long* c, int it) { mix<2><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    run("mix + fadd2 + int max", [=](int g, float* o, long long* c, int it) { mix<3><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    run("phased: +fadd2 +int max", [=](int g, float* o, long long* c, int it) { mixp<3><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    run("phased: +fadd2", [=](int g, float* o, long long* c, int it) { mixp<2><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    run("phased: +fadd2 +sumcheck", [=](int g, float* o, long long* c, int it) { mixp<6><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
#define M2(S, C, NAME) run(NAME, [=](int g, float* o, long long* c, int it) { mix2<S, C><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
    M2(0, 0, "mix2 max, no sum")
    M2(1, 0, "mix2 FADD2 x2acc (current)")
    M2(2, 0, "mix2 FADD2 x4acc")
    M2(3, 0, "mix2 scalar FADD x4acc")
    M2(4, 0, "mix2 HADD2 fp16 accumulate")
    M2(5, 0, "mix2 HADD2 widen every 8 words")
    M2(0, 1, "mix2 no sum, magic-add cvt")
    M2(1, 1, "mix2 FADD2 x2acc, magic-add cvt")
    M2(4, 1, "mix2 HADD2, magic-add cvt")
    run("phased: no sum", [=](int g, float* o, long long* c, int it) { mixp<0><<<g, 128>>>(o, c, it, 1e-4f, in); }, bps, 32, 2000);
  }
  return 0;
}
