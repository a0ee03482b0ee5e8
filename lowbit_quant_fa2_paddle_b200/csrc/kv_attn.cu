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
//
// Both contractions walk the packed words as they are: one 32-bit word holds 8 (4-bit) or 16 (2-bit) codes that share
// their scale group, so a thread takes 8 codes per shared-memory load and the affine part is factored out,
//   S[n]  = sum_d (q[d] sc[d,g]) code[d,n]  +  sum_d q[d] mn[d,g]          (g = group of key n)
//   o[c]  = sum_n (p[n] vs[n,g]) code[n,c]  +  sum_n p[n] vm[n,g]          (g = group of channel c)
// which leaves one conversion and one FMA per code in the inner loops (the first version dequantized every element
// with its own loads: 9 instructions per code, 1.0 TB/s).
//   scores : thread (ko, ds) <-> key octet ko of the tile x channel slice {d = 8 j + ds}; the 8 slices of an octet are
//            adjacent lanes and meet by three xor-shuffles
//   P.V    : thread (co, ks) <-> channel octet co x key slice ks; the slices meet once per CTA, at the end
template <int D, int BITS, int R>
__global__ void __launch_bounds__(kKvThreads)
kv_attn_partial_kernel(const __half* __restrict__ q, const uint8_t* __restrict__ kcode,
                       const __half* __restrict__ kscale, const __half* __restrict__ kmn,
                       const uint8_t* __restrict__ vcode, const __half* __restrict__ vscale,
                       const __half* __restrict__ vmn, float* __restrict__ ws, int H, int Nq, int N, int nsplit,
                       int keys_per_split, float scale_log2e, int64_t qsb, int64_t qsn, int64_t qsh) {
  constexpr int VB = D * BITS / 8;              // bytes of one token row of V
  constexpr uint32_t CM = (1u << BITS) - 1u;
  constexpr int KG = kKvTile / kKvGroup;        // K scale groups per tile (4)
  constexpr int VG = D / kKvGroup;              // V scale groups per token
  constexpr int KW = kKvTile * BITS / 32;       // words per channel row of a K tile
  constexpr int VW = D * BITS / 32;             // words per token row of V
  constexpr int KWP = KW + 4, VWP = VW + 1;     // padded row strides: the access patterns below are conflict-free
  constexpr int DS = D / 8;                     // channels per score thread
  constexpr int NCO = D / 8;                    // channel octets
  constexpr int KPS = kKvTile / (kKvThreads / NCO);  // keys per P.V thread and tile
  static_assert(kKvThreads == 128 && kKvTile == 128, "thread <-> (octet, slice) maps assume 128 x 128");
  __shared__ __align__(16) float sQ[R][D];
  __shared__ uint32_t sK[D][KWP];
  __shared__ float2 sKs[D][KG];                 // (scale, mn)
  __shared__ uint32_t sV[kKvTile][VWP];
  __shared__ float2 sVs[kKvTile][VG];           // (scale, mn)
  __shared__ float2 sPV[R][kKvTile][VG];        // (p * scale, p * mn); reused for the final cross-warp fold
  __shared__ float sRedM[R][4], sRedL[R][4];

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

  // thread maps
  const int ko = tid >> 3, ds = tid & 7;        // scores: key octet (a warp = 4 octets = one scale group), channel slice
  const int kwidx = (ko * 8 * BITS) / 32, kwsh = (ko * 8 * BITS) % 32;
  const int co = tid % NCO, ks = tid / NCO;     // P.V: channel octet, key slice
  const int vwidx = (co * 8 * BITS) / 32, vwsh = (co * 8 * BITS) % 32;
  const int gch = co / 4;                       // scale group of my channels (32 channels = 4 octets)

  float m_run[R], l_run[R], o_acc[R][8];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) o_acc[r][i] = 0.f;
  }

  // ---- tile staging, split in two halves: `fetch` issues the global loads of a tile into registers (4-byte words,
  //      consecutive threads on consecutive words of a row, every thread keeps its column and walks the rows with a
  //      constant pointer stride), `commit` parks them in shared memory.  The loads of tile i+1 are issued before
  //      the arithmetic of tile i, so their latency hides behind it. ----
  constexpr int KRS = kKvThreads / KW, NKW = D / KRS;        // K rows per pass, K words per thread
  constexpr int SRS = kKvThreads / KG, NKS = D / SRS;        // K scale rows per pass, entries per thread
  constexpr int VRS = kKvThreads / VW, NVW = kKvTile / VRS;  // V rows per pass, V words per thread
  constexpr int TRS = kKvThreads / VG, NVS = kKvTile / TRS;  // V scale rows per pass, entries per thread
  const int kw = tid % KW, kd0 = tid / KW;
  const int sg = tid % KG, sd0 = tid / KG;
  const int vw = tid % VW, vn0 = tid / VW;
  const int tg = tid % VG, tn0 = tid / VG;
  uint32_t rk[NKW], rv[NVW];
  __half rks[NKS], rkm[NKS], rvs[NVS], rvm[NVS];
  auto fetch = [&](int n0) {
    const bool k_ok = (n0 + kw * (32 / BITS)) < N;  // N is a multiple of 32: a word never straddles the end
    const uint8_t* kp = kc_base + (int64_t)kd0 * H * krow_bytes + (int64_t)n0 * BITS / 8 + kw * 4;
    const int64_t kstep = (int64_t)KRS * H * krow_bytes;
#pragma unroll
    for (int i = 0; i < NKW; ++i, kp += kstep) rk[i] = k_ok ? *reinterpret_cast<const uint32_t*>(kp) : 0u;
    const int gi = n0 / kKvGroup + sg;
    const bool s_ok = gi < kgroups;
    const __half* sp = ks_base + (int64_t)sd0 * H * kgroups + gi;
    const __half* mp = km_base + (int64_t)sd0 * H * kgroups + gi;
    const int64_t sstep = (int64_t)SRS * H * kgroups;
#pragma unroll
    for (int i = 0; i < NKS; ++i, sp += sstep, mp += sstep) {
      rks[i] = s_ok ? *sp : __float2half_rn(0.f);
      rkm[i] = s_ok ? *mp : __float2half_rn(0.f);
    }
    const uint8_t* vp = vc_base + (int64_t)(n0 + vn0) * H * VB + vw * 4;
    const int64_t vstep = (int64_t)VRS * H * VB;
#pragma unroll
    for (int i = 0; i < NVW; ++i, vp += vstep) rv[i] = (n0 + vn0 + i * VRS < N) ? *reinterpret_cast<const uint32_t*>(vp) : 0u;
    const __half* tsp = vs_base + (int64_t)(n0 + tn0) * H * VG + tg;
    const __half* tmp_ = vm_base + (int64_t)(n0 + tn0) * H * VG + tg;
    const int64_t tstep = (int64_t)TRS * H * VG;
#pragma unroll
    for (int i = 0; i < NVS; ++i, tsp += tstep, tmp_ += tstep) {
      const bool ok = (n0 + tn0 + i * TRS) < N;
      rvs[i] = ok ? *tsp : __float2half_rn(0.f);
      rvm[i] = ok ? *tmp_ : __float2half_rn(0.f);
    }
  };
  auto commit = [&]() {
#pragma unroll
    for (int i = 0; i < NKW; ++i) sK[kd0 + i * KRS][kw] = rk[i];
#pragma unroll
    for (int i = 0; i < NKS; ++i) sKs[sd0 + i * SRS][sg] = make_float2(__half2float(rks[i]), __half2float(rkm[i]));
#pragma unroll
    for (int i = 0; i < NVW; ++i) sV[vn0 + i * VRS][vw] = rv[i];
#pragma unroll
    for (int i = 0; i < NVS; ++i) sVs[tn0 + i * TRS][tg] = make_float2(__half2float(rvs[i]), __half2float(rvm[i]));
  };

  fetch(n_begin);
  for (int n0 = n_begin; n0 < n_end; n0 += kKvTile) {
    __syncthreads();  // previous tile fully consumed (also orders the sQ fill before its first use)
    commit();
    __syncthreads();
    if (n0 + kKvTile < n_end) fetch(n0 + kKvTile);  // in flight during this tile's arithmetic

    // ---- scores of my key octet over my channel slice, then the 8 slices meet ----
    float s[R][8], cst[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cst[r] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[r][i] = 0.f;
    }
#pragma unroll 2
    for (int j = 0; j < DS; ++j) {
      const int d = j * 8 + ds;
      const uint32_t codes = sK[d][kwidx] >> kwsh;
      const float2 sm = sKs[d][warp];  // the warp's keys share one scale group
      float qs[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float qv = sQ[r][d];
        qs[r] = qv * sm.x;
        cst[r] = fmaf(qv, sm.y, cst[r]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float cf = (float)((codes >> (i * BITS)) & CM);
#pragma unroll
        for (int r = 0; r < R; ++r) s[r][i] = fmaf(qs[r], cf, s[r][i]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        cst[r] += __shfl_xor_sync(0xffffffffu, cst[r], off);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[r][i] += __shfl_xor_sync(0xffffffffu, s[r][i], off);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s[r][i] += cst[r];
    }
    // ---- online softmax (base 2); every lane of an octet holds the same 8 scores ----
    const int nk = n0 + ko * 8;  // first key of my octet
    float mloc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) v = fmaxf(v, (nk + i < n_end) ? s[r][i] : -INFINITY);
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
      mloc[r] = v;
      if (lane == 0) sRedM[r][warp] = v;
    }
    __syncthreads();
    float alpha[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mt = fmaxf(fmaxf(sRedM[r][0], sRedM[r][1]), fmaxf(sRedM[r][2], sRedM[r][3]));
      const float m_new = fmaxf(m_run[r], mt);   // finite: every tile holds at least one live key
      alpha[r] = kv_ex2(m_run[r] - m_new);       // first tile: exp2(-inf) = 0
      m_run[r] = m_new;
      float lsum = 0.f, pmine = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float pi = (nk + i < n_end) ? kv_ex2(s[r][i] - m_new) : 0.f;
        lsum += pi;
        pmine = (i == ds) ? pi : pmine;  // lane ds of the octet publishes key ds
      }
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
      if (lane == 0) sRedL[r][warp] = lsum;
      const int nl = ko * 8 + ds;
#pragma unroll
      for (int g = 0; g < VG; ++g) {
        const float2 vsm = sVs[nl][g];
        sPV[r][nl][g] = make_float2(pmine * vsm.x, pmine * vsm.y);
      }
    }
    (void)mloc;
    __syncthreads();
    // ---- P.V over my channel octet and key slice ----
#pragma unroll
    for (int r = 0; r < R; ++r) {
      l_run[r] = l_run[r] * alpha[r] + ((sRedL[r][0] + sRedL[r][1]) + (sRedL[r][2] + sRedL[r][3]));
#pragma unroll
      for (int i = 0; i < 8; ++i) o_acc[r][i] *= alpha[r];
    }
    {
      float cv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) cv[r] = 0.f;
#pragma unroll 2
      for (int nn = 0; nn < KPS; ++nn) {
        const int nl = ks * KPS + nn;
        const uint32_t codes = sV[nl][vwidx] >> vwsh;
        float pv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 t = sPV[r][nl][gch];
          pv[r] = t.x;
          cv[r] += t.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float cf = (float)((codes >> (i * BITS)) & CM);
#pragma unroll
          for (int r = 0; r < R; ++r) o_acc[r][i] = fmaf(pv[r], cf, o_acc[r][i]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o_acc[r][i] += cv[r];
      }
    }
  }

  // ---- partial result of this split: fold the key slices (lanes, then warps) ----
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int off = NCO; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o_acc[r][i] += __shfl_xor_sync(0xffffffffu, o_acc[r][i], off);
    }
  }
  __syncthreads();  // sPV is free
  float* sO = reinterpret_cast<float*>(&sPV[0][0][0]);  // [4 warps][R][D]
  if (lane < NCO) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sO[(warp * R + r) * D + co * 8 + i] = o_acc[r][i];
    }
  }
  __syncthreads();
  for (int i = tid; i < R * D; i += kKvThreads) {
    const int r = i / D, c = i % D;
    if (row0 + r >= Nq) continue;
    const float v = (sO[(0 * R + r) * D + c] + sO[(1 * R + r) * D + c]) + (sO[(2 * R + r) * D + c] + sO[(3 * R + r) * D + c]);
    float* dst = ws + ((((int64_t)b * H + h) * Nq + (row0 + r)) * nsplit + split) * (D + 2);
    dst[2 + c] = v;
  }
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= Nq) continue;
      float* dst = ws + ((((int64_t)b * H + h) * Nq + (row0 + r)) * nsplit + split) * (D + 2);
      dst[0] = m_run[r];
      dst[1] = l_run[r];
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
  // splits of whole tiles, none of them empty; about four waves of the resident capacity (148 SMs x 6-7 CTAs) so
  // that the partial last wave costs little (the first version's 10 splits made 1.24 waves: 38 % of the run at a
  // quarter of the occupancy)
  const int tiles = (N + kKvTile - 1) / kKvTile;
  const int64_t base = (int64_t)B * H * ((Nq + R - 1) / R);
  int want = (int)((148 * 6 * 4) / base);
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
