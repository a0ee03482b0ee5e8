// kv_attn.cu -- attention over a KIVI-packed low-bit K/V cache (fp16 queries), SURVEY 8(f) rank 4.
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   _quantized_flash_attn_forward + _fwd_kernel   src/triton/quantization/attn_4bit_per_block.py:28-553
//     (the prototype's driver, :637-690, feeds it new_pack.py:247-300 codes: K packed along the sequence per channel,
//      V packed along the channels per token, asymmetric, group 32, scale + minimum)
// Arithmetic that defines the result (attn_4bit_per_block.py:260-262, 330-372; new_pack.py:69-144):
//   K^[d,n] = fma(code, scale[d, n/G], mn[d, n/G])      V^[n,d] = fma(code, scale[n, d/G], mn[n, d/G])     (fp32)
//   S = q . K^ (fp32),  p = exp(S * softmax_scale - m),  o = (sum_n p V^) / l,  lse = m + log(l)
// The reference kernel is a prototype that does not run as written, so parity for this entry is
// unpinned: the tests' CPU restatement follows the formulas above over the reference's own (pinned) pack format.
//
// This is a decode-shaped, HBM-bound path (a few query rows against a long packed cache), not a tensor-core one:
// the cache is read exactly once.  Grid (key splits x query-row tiles, heads, batch); a CTA of 128 threads walks
// its key range in 128-key tiles staged in shared memory:
//   scores : thread t <-> key t of the tile (a warp = one 32-key scale group of K), loop over channels
//   softmax: online, base 2, row maximum / sum by warp shuffles + one shared-memory hop
//   P.V    : thread c <-> channel c (a warp = one 32-channel scale group of V), loop over the tile's keys
// and leaves (m, l, o) of its split in a workspace; kv_attn_merge_kernel folds the splits (flash-decoding).
#include "common.cuh"

namespace lowbit {

constexpr int kKvThreads = 128;
constexpr int kKvTile = 128;   // keys per tile
constexpr int kKvGroup = 32;   // quantization group (new_pack.py driver: group_size=32)

__device__ __forceinline__ float kv_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// R query rows per CTA.  Workspace layout per (b, h, row, split): [m, l, o[D]] fp32.
template <int D, int BITS, int R>
__global__ void __launch_bounds__(kKvThreads)
kv_attn_partial_kernel(const __half* __restrict__ q, const uint8_t* __restrict__ kcode,
                       const __half* __restrict__ kscale, const __half* __restrict__ kmn,
                       const uint8_t* __restrict__ vcode, const __half* __restrict__ vscale,
                       const __half* __restrict__ vmn, float* __restrict__ ws, int H, int Nq, int N, int nsplit,
                       int keys_per_split, float scale_log2e, int64_t qsb, int64_t qsn, int64_t qsh) {
  constexpr int KB = kKvTile * BITS / 8;        // bytes of one channel row of a K tile
  constexpr int VB = D * BITS / 8;              // bytes of one token row of V
  constexpr int CPB = 8 / BITS;                 // codes per byte
  constexpr uint32_t CM = (1u << BITS) - 1u;
  constexpr int KG = kKvTile / kKvGroup;        // K scale groups per tile (4)
  constexpr int VG = D / kKvGroup;              // V scale groups per token
  constexpr int NH = kKvThreads / D;            // key halves in the P.V phase (D = 64: 2, D = 128: 1)
  __shared__ __align__(16) float sQ[R][D];
  __shared__ __align__(16) uint8_t sK[D][KB];
  __shared__ __half2 sKs[D][KG];                // (scale, mn)
  __shared__ __align__(16) uint8_t sV[kKvTile][VB];
  __shared__ __half2 sVs[kKvTile][VG];
  __shared__ float sP[R][kKvTile];
  __shared__ float sRed[R][4];
  __shared__ float sO[NH > 1 ? R : 1][NH > 1 ? D : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x % nsplit, qtile = blockIdx.x / nsplit;
  const int h = blockIdx.y, b = blockIdx.z;
  const int row0 = qtile * R;
  const int n_begin = split * keys_per_split, n_end = min(N, n_begin + keys_per_split);

  // queries of this tile, pre-multiplied by softmax_scale * log2(e) (the softmax below runs in base 2)
  for (int i = tid; i < R * D; i += kKvThreads) {
    const int r = i / D, d = i % D;
    float v = 0.f;
    if (row0 + r < Nq) v = __half2float(q[b * qsb + (int64_t)(row0 + r) * qsn + h * qsh + d]) * scale_log2e;
    sQ[r][d] = v;
  }

  // cache addressing: kcode [B][D][H][N*BITS/8], kscale/kmn [B][D][H][N/G]; vcode [B][N][H][VB], vscale/vmn [B][N][H][VG]
  const int64_t krow_bytes = (int64_t)N * BITS / 8, kgroups = N / kKvGroup;
  const uint8_t* kc_base = kcode + ((int64_t)b * D * H + h) * krow_bytes;     // + d * H * krow_bytes
  const __half* ks_base = kscale + ((int64_t)b * D * H + h) * kgroups;
  const __half* km_base = kmn + ((int64_t)b * D * H + h) * kgroups;
  const uint8_t* vc_base = vcode + ((int64_t)b * N * H + h) * VB;            // + n * H * VB
  const __half* vs_base = vscale + ((int64_t)b * N * H + h) * VG;
  const __half* vm_base = vmn + ((int64_t)b * N * H + h) * VG;

  float m_run[R], l_run[R], o_acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { m_run[r] = -INFINITY; l_run[r] = 0.f; o_acc[r] = 0.f; }
  const int c = tid % D;         // my channel in the P.V phase
  const int khalf = tid / D;     // my half of the tile's keys in the P.V phase (D = 64)

  for (int n0 = n_begin; n0 < n_end; n0 += kKvTile) {
    __syncthreads();  // previous tile fully consumed (also orders the sQ fill before its first use)
    // ---- stage the tile: 4-byte words, consecutive threads on consecutive words of a row ----
    {
      constexpr int KW = KB / 4;  // words per channel row of the K tile
      for (int i = tid; i < D * KW; i += kKvThreads) {
        const int d = i / KW, w = i % KW;
        const int n = n0 + w * 4 * CPB;  // first key of this word
        uint32_t v = 0;
        if (n < N) v = *reinterpret_cast<const uint32_t*>(kc_base + (int64_t)d * H * krow_bytes + (int64_t)n0 * BITS / 8 + w * 4);
        *reinterpret_cast<uint32_t*>(&sK[d][w * 4]) = v;
      }
      for (int i = tid; i < D * KG; i += kKvThreads) {
        const int d = i / KG, g = i % KG;
        const int gi = n0 / kKvGroup + g;
        __half2 v = __floats2half2_rn(0.f, 0.f);
        if (gi < kgroups) v = __halves2half2(ks_base[(int64_t)d * H * kgroups + gi], km_base[(int64_t)d * H * kgroups + gi]);
        sKs[d][g] = v;
      }
      constexpr int VW = VB / 4;  // words per token row of V
      for (int i = tid; i < kKvTile * VW; i += kKvThreads) {
        const int nl = i / VW, w = i % VW;
        uint32_t v = 0;
        if (n0 + nl < N) v = *reinterpret_cast<const uint32_t*>(vc_base + (int64_t)(n0 + nl) * H * VB + w * 4);
        *reinterpret_cast<uint32_t*>(&sV[nl][w * 4]) = v;
      }
      for (int i = tid; i < kKvTile * VG; i += kKvThreads) {
        const int nl = i / VG, g = i % VG;
        __half2 v = __floats2half2_rn(0.f, 0.f);
        if (n0 + nl < N) v = __halves2half2(vs_base[(int64_t)(n0 + nl) * H * VG + g], vm_base[(int64_t)(n0 + nl) * H * VG + g]);
        sVs[nl][g] = v;
      }
    }
    __syncthreads();

    // ---- scores: thread <-> key tid of the tile ----
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = 0.f;
    {
      const int byte = tid / CPB, sh = (tid % CPB) * BITS, g = tid / kKvGroup;
#pragma unroll 4
      for (int d = 0; d < D; ++d) {
        const float code = (float)((sK[d][byte] >> sh) & CM);
        const float2 sm = __half22float2(sKs[d][g]);
        const float kd = fmaf(code, sm.x, sm.y);
#pragma unroll
        for (int r = 0; r < R; ++r) s[r] = fmaf(sQ[r][d], kd, s[r]);
      }
    }
    const bool live = (n0 + tid) < n_end;
    // ---- online softmax (base 2) ----
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = live ? s[r] : -INFINITY;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
      if (lane == 0) sRed[r][warp] = v;
    }
    __syncthreads();
    float alpha[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mt = fmaxf(fmaxf(sRed[r][0], sRed[r][1]), fmaxf(sRed[r][2], sRed[r][3]));
      const float m_new = fmaxf(m_run[r], mt);   // finite: every tile holds at least one live key
      alpha[r] = kv_ex2(m_run[r] - m_new);       // first tile: exp2(-inf) = 0
      p[r] = live ? kv_ex2(s[r] - m_new) : 0.f;
      m_run[r] = m_new;
      sP[r][tid] = p[r];
    }
    __syncthreads();  // sRed read by everyone before it is reused for the sums; sP complete
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = p[r];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) sRed[r][warp] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      l_run[r] = l_run[r] * alpha[r] + ((sRed[r][0] + sRed[r][1]) + (sRed[r][2] + sRed[r][3]));
      o_acc[r] *= alpha[r];
    }
    // ---- P.V: thread <-> channel c (and key half khalf when D = 64) ----
    {
      const int byte = c / CPB, sh = (c % CPB) * BITS, g = c / kKvGroup;
      constexpr int KPH = kKvTile / NH;  // keys per half
      const int nb = khalf * KPH;
#pragma unroll 4
      for (int nl = nb; nl < nb + KPH; ++nl) {
        const float code = (float)((sV[nl][byte] >> sh) & CM);
        const float2 sm = __half22float2(sVs[nl][g]);
        const float vd = fmaf(code, sm.x, sm.y);
#pragma unroll
        for (int r = 0; r < R; ++r) o_acc[r] = fmaf(sP[r][nl], vd, o_acc[r]);
      }
    }
  }

  // ---- partial result of this split ----
  if constexpr (NH > 1) {  // fold the two key halves (same channel, threads c and c + D)
    __syncthreads();
    if (khalf == 1) {
#pragma unroll
      for (int r = 0; r < R; ++r) sO[r][c] = o_acc[r];
    }
    __syncthreads();
    if (khalf == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) o_acc[r] += sO[r][c];
    }
  }
  if (khalf == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= Nq) continue;
      float* dst = ws + ((((int64_t)b * H + h) * Nq + (row0 + r)) * nsplit + split) * (D + 2);
      if (c == 0) { dst[0] = m_run[r]; dst[1] = l_run[r]; }
      dst[2 + c] = o_acc[r];
    }
  }
}

// o = sum_s o_s 2^(m_s - M) / sum_s l_s 2^(m_s - M),  lse = ln2 * (M + log2 L)   (natural log, like the reference's lse)
template <int D>
__global__ void kv_attn_merge_kernel(const float* __restrict__ ws, __half* __restrict__ o, float* __restrict__ lse,
                                     int H, int Nq, int nsplit, int64_t osb, int64_t osn, int64_t osh, int lse_stride) {
  const int row = blockIdx.x, h = blockIdx.y, b = blockIdx.z, c = threadIdx.x;  // D threads
  const float* src = ws + (((int64_t)b * H + h) * Nq + row) * nsplit * (D + 2);
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, src[(int64_t)s * (D + 2)]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float w = kv_ex2(src[(int64_t)s * (D + 2)] - M);
    L += src[(int64_t)s * (D + 2) + 1] * w;
    acc += src[(int64_t)s * (D + 2) + 2 + c] * w;
  }
  o[b * osb + (int64_t)row * osn + h * osh + c] = __float2half_rn(acc / L);
  if (c == 0 && lse != nullptr) lse[((int64_t)b * H + h) * lse_stride + row] = 0.6931471805599453f * (M + log2f(L));
}

static int kv_splits(int B, int H, int Nq, int N, int R, int* keys_per_split) {
  // enough CTAs to fill the GPU a few times over, splits of whole tiles, none of them empty
  const int tiles = (N + kKvTile - 1) / kKvTile;
  const int64_t base = (int64_t)B * H * ((Nq + R - 1) / R);
  int want = (int)((148 * 8 + base - 1) / base);
  want = want < 1 ? 1 : (want > tiles ? tiles : want);
  const int tps = (tiles + want - 1) / want;  // tiles per split
  *keys_per_split = tps * kKvTile;
  return (tiles + tps - 1) / tps;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" int64_t lowbit_kv_attn_workspace_bytes(int B, int H, int Nq, int N, int D) {
  int kps = 0;
  const int R = Nq > 1 ? 4 : 1;
  const int ns = kv_splits(B, H, Nq, N, R, &kps);
  return (int64_t)B * H * Nq * ns * (D + 2) * 4;
}

extern "C" int lowbit_kv_attn_fwd(const void* q, const void* kcode, const void* kscale, const void* kmn,
                                  const void* vcode, const void* vscale, const void* vmn, void* o, float* lse,
                                  void* workspace, int B, int H, int Nq, int N, int D, int group_size, int bits,
                                  float softmax_scale, int64_t qsb, int64_t qsn, int64_t qsh, int64_t osb, int64_t osn,
                                  int64_t osh, int lse_stride, void* stream) {
  LOWBIT_CHECK(q && kcode && kscale && kmn && vcode && vscale && vmn && o && workspace, "lowbit_kv_attn_fwd: null pointer");
  LOWBIT_CHECK(D == 64 || D == 128, "lowbit_kv_attn_fwd: head_dim must be 64 or 128 (got %d)", D);
  LOWBIT_CHECK(bits == 4 || bits == 2, "lowbit_kv_attn_fwd: bits must be 4 or 2 (got %d)", bits);
  LOWBIT_CHECK(group_size == kKvGroup, "lowbit_kv_attn_fwd: group_size must be 32 (got %d)", group_size);
  LOWBIT_CHECK(B > 0 && H > 0 && Nq > 0 && N > 0, "lowbit_kv_attn_fwd: empty tensor");
  LOWBIT_CHECK(N % kKvGroup == 0, "lowbit_kv_attn_fwd: the cache length must be a multiple of the group size (got %d)", N);
  LOWBIT_CHECK(((uintptr_t)kcode & 3) == 0 && ((uintptr_t)vcode & 3) == 0, "lowbit_kv_attn_fwd: codes must be 4-byte aligned");
  LOWBIT_CHECK(lse == nullptr || lse_stride >= Nq, "lowbit_kv_attn_fwd: lse_stride < Nq");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = Nq > 1 ? 4 : 1;
  int kps = 0;
  const int ns = kv_splits(B, H, Nq, N, R, &kps);
  const int qtiles = (Nq + R - 1) / R;
  dim3 grid((unsigned)(ns * qtiles), H, B);
  const float sl2 = softmax_scale * 1.4426950408889634f;
#define KV_LAUNCH(DD, BB, RR)                                                                                          \
  kv_attn_partial_kernel<DD, BB, RR><<<grid, kKvThreads, 0, st>>>(                                                     \
      (const __half*)q, (const uint8_t*)kcode, (const __half*)kscale, (const __half*)kmn, (const uint8_t*)vcode,       \
      (const __half*)vscale, (const __half*)vmn, (float*)workspace, H, Nq, N, ns, kps, sl2, qsb, qsn, qsh)
#define KV_BY_R(DD, BB) do { if (R == 1) KV_LAUNCH(DD, BB, 1); else KV_LAUNCH(DD, BB, 4); } while (0)
  if (D == 64) { if (bits == 4) KV_BY_R(64, 4); else KV_BY_R(64, 2); }
  else { if (bits == 4) KV_BY_R(128, 4); else KV_BY_R(128, 2); }
#undef KV_BY_R
#undef KV_LAUNCH
  LOWBIT_CUDA(cudaGetLastError());
  dim3 mgrid(Nq, H, B);
  if (D == 64)
    kv_attn_merge_kernel<64><<<mgrid, 64, 0, st>>>((const float*)workspace, (__half*)o, lse, H, Nq, ns, osb, osn, osh, lse_stride);
  else
    kv_attn_merge_kernel<128><<<mgrid, 128, 0, st>>>((const float*)workspace, (__half*)o, lse, H, Nq, ns, osb, osn, osh, lse_stride);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}
