// ubench_softmax.cu -- instruction-mix microbenchmark of the softmax step of csrc/attn.cu (attn_fwd_n64_kernel): the
// SAME device functions (csrc/softmax_chunk.cuh), no TMEM / barriers, 128 scores per iteration, 8 warps per SM
// (2 CTAs of 128 threads, shared memory sized so that exactly two fit).  Development aid, not part of the library.
// It gives the ceiling the instruction mix alone allows (profiles/r2_ubench_softmax.log): exact maximum + all MUFU
// 13.9 exp2/clk/SM, optimistic maximum 15.5, optimistic + 1/8 on the FMA pipe 15.9 (MUFU alone: 16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_softmax tools/ubench_softmax.cu
// Prints exp2 results per clock per SM (peak of the MUFU pipe alone: 16).
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>
#include <cstdio>

#include "../lowbit_quant_fa2_paddle_b200/csrc/softmax_chunk.cuh"

using namespace lowbit::chunk;

// EXACT = 1: integer row max over the 128 scores first (the first / masked step of a row)
// EXACT = 0: optimistic step (stale maximum, overflow check on the row sum)
template <int PF, int EXACT>
__global__ void __maxnreg__(200) step_kernel(float* out, long long* cyc, int iters, float sc, const int* in) {
  uint32_t s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = (uint32_t)(in[i] + (int)threadIdx.x);
  float l = 0.f, m_ref = 0.f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (EXACT) {
      const int ia = row_max_i<64, false>(s, 0), ib = row_max_i<64, false>(s + 64, 0);
      const float mb = fmaxf((float)ia * sc, (float)ib * sc);
      if (__any_sync(0xffffffffu, mb > m_ref + 8.f)) m_ref = fmaxf(m_ref, mb);
    }
    const float nm = -m_ref;
    float lsum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
      lsum += chunk_f16<false, PF>(s + 32 * c, sc, nm, 0, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= pk[i];
    }
    if (!EXACT && __any_sync(0xffffffffu, !(lsum < 32768.f))) m_ref += 1.f;
    l += lsum;
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] += (acc & 1);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(acc) + m_ref;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PF, int EXACT>
static void run(const char* name, float* out, long long* cyc, const int* in, int sms) {
  const int iters = 2000, grid = sms * 2;
  auto kern = step_kernel<PF, EXACT>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  kern<<<grid, 128, 100 * 1024>>>(out, cyc, 10, 1e-4f, in);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<<<grid, 128, 100 * 1024>>>(out, cyc, iters, 1e-4f, in);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  long long h[1024];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += (double)h[i];
  mean /= grid;
  const double exps_per_sm = 2.0 * 128 * 128.0 * iters;  // 2 CTAs x 128 threads x 128 scores
  printf("%-44s %6.2f exp2/clk/SM (SM clocks)  %6.2f exp2/clk/SM @1965MHz  (%.3f ms, %s)\n", name, exps_per_sm / mean,
         exps_per_sm / (ms * 1e-3 * 1965e6), ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  float* out;
  long long* cyc;
  int* in;
  cudaMalloc(&out, sizeof(float) * sms * 2 * 128);
  cudaMalloc(&cyc, sizeof(long long) * sms * 2);
  cudaMalloc(&in, sizeof(int) * 128);
  int hin[128];
  for (int i = 0; i < 128; ++i) hin[i] = -(i * 977 % 50000);
  cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
  run<0, 1>("exact max, all MUFU", out, cyc, in, sms);
  run<0, 0>("optimistic, all MUFU", out, cyc, in, sms);
  run<1, 0>("optimistic, 1/8 on the FMA pipe", out, cyc, in, sms);
  run<2, 0>("optimistic, 2/8 on the FMA pipe", out, cyc, in, sms);
  run<3, 0>("optimistic, 3/8 on the FMA pipe", out, cyc, in, sms);
  run<4, 0>("optimistic, 4/8 on the FMA pipe", out, cyc, in, sms);
  run<2, 1>("exact max, 2/8 on the FMA pipe", out, cyc, in, sms);
  run<8, 0>("optimistic, all on the FMA pipe", out, cyc, in, sms);
  return 0;
}
