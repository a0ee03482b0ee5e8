// kv_attn.cu -- attention over a KIVI-packed low-bit K/V cache (fp16 queries), SURVEY 8(f) rank 4.
//
// Replaces, behind include/lowbit_fa.h (paths relative to the reference repository):
//   _quantized_flash_attn_forward + _fwd_kernel   src/triton/quantization/attn_4bit_per_block.py:28-553
//     (the prototype's driver, :637-690, feeds it new_pack.py:247-300 codes: K packed along the sequence per channel,
//      V packed along the channels per token, asymmetric, group 32, scale + minimum)
// Arithmetic that defines the result (attn_4bit_per_block.py:260-262, 330-372; new_pack.py:69-144):
//   K^[d,n] = fma(code, scale[d, n/G], mn[d, n/G])      V^[n,d] = fma(code, scale[n, d/G], mn[n, d/G])
//   S = q . K^,  p = exp(S * softmax_scale - m),  o = (sum_n p V^) / l,  lse = m + log(l)
// The reference kernel is a prototype that does not run as written; parity is pinned to what its
// own driver checks it against (:637-788): attention over the caches dequantized by the reference's unpack_and_dequant_*
// functions, run unmodified (tests/golden/kvcache_*.npz, tools/make_golden_kvcache.py), on the pinned pack format.
//
// A decode-shaped, HBM-bound path: up to 8 query rows against a long packed cache that is read exactly once.  The
// first version dequantized every code with shift / mask / I2F / FMA (14.8 thread instructions per code, 27 % of
// the HBM roofline, issue bound).  This one never converts a code:
//
//  * both contractions run on mma.sync.m16n8k16 (fp16 in, fp32 accumulate) with the CODES as the A operand
//    (m = 16 keys resp. 16 channels, k = 16 channels resp. 16 keys) and the query side as B (n = 8 query rows);
//  * a packed 32-bit word (8 4-bit or 16 2-bit codes of ONE channel resp. token) and its neighbour along k are merged
//    by one PRMT into "lo half = 16 code bits of k, hi half = 16 code bits of k + 1", and every A register is then one
//    LOP3 of that word with a mask: the bits land in the mantissa of an fp16 pair.  Read as fp16 DENORMALS they are
//    code * f * 2^-24 with f = 2^(bit position) -- exact -- and the power of two leaves through one fp32 multiply per
//    score / output element.  (MAGIC variant: OR in 0x6400 and subtract 1024 in fp16, for hardware that would flush);
//  * the affine parts are factored out as before:  S[n] = sum_d (q[d] sc[d,g]) code[d,n] + sum_d q[d] mn[d,g]; the
//    second sum is one more MMA against a tile of ones (all rows equal), likewise for P.V;
//  * the B operands are fp16 products q * sc, q * mn, p * vs, p * vm (one HMUL2 per register; scales staged
//    TRANSPOSED in shared memory so that a register's two k-neighbours are one 32-bit load).
//
// Grid (key splits x query-row tiles, heads, batch); CTA = 4 warps; a 128-key tile is staged by 16-byte cp.async
// into a 3-deep ring (codes exactly as they lie in HBM, rows padded against bank conflicts); warp w owns keys
// [32 w, 32 w + 32) of every tile for BOTH contractions with its own online-softmax state, so the main loop has one
// __syncthreads per tile (ring hand-over) and no cross-warp exchange; the four states are merged once per CTA and the
// split's (m, l, o) goes to the workspace that kv_attn_merge_kernel folds (flash-decoding).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace lowbit {

constexpr int kKvWarps = 8;
constexpr int kKvThreads = 32 * kKvWarps;
constexpr int kKvTile = 32 * kKvWarps;   // keys per tile: one 32-key scale group per warp; a K row of a tile is a whole 128-byte line (4-bit)
constexpr int kKvGroup = 32;   // quantization group (new_pack.py driver: group_size=32)
constexpr int kKvRows = 8;     // query rows per CTA: n of the m16n8k16 MMA
constexpr int kKvStages = 2;
constexpr int kKvCtasPerSm = 2;
constexpr int kKvPStride = 40; // halves per row of the per-warp P tile (32 keys + pad: conflict-free both ways)

__device__ __forceinline__ float kv_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int BYTES>
__device__ __forceinline__ void kv_cp_async(uint32_t dst, const void* src, int valid) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
  else if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void kv_cp_async_full(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void kv_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void kv_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// D(16x8, fp32) += A(16x16, fp16, row) . B(16x8, fp16, col)
__device__ __forceinline__ void kv_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t kv_hmul2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t kv_pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// shared-memory loads at (32-bit shared address + compile-time offset): the offset lands in the instruction's immediate
template <int OFF> __device__ __forceinline__ uint32_t kv_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF> __device__ __forceinline__ uint4 kv_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a), "n"(OFF) : "memory");
  return v;
}

// Four A registers (fp16 pairs: lo half <- the first word's code, hi half <- the second word's) out of x, the PRMT
// merge of two packed words.  A lane owns 4 consecutive codes of each word: 4-bit -- the 16 bits ARE its codes, at
// bit 0, 4, 8, 12 (the upper two are brought down by 8 so that they stay inside the 10 mantissa bits); 2-bit -- the 16
// bits hold 8 codes, the lane's four start at bit 8 * quad.  Code i comes out multiplied by kFactor[i] (* 2^-24).
template <int BITS, bool MAGIC>
struct KvDec {
  static __device__ __forceinline__ uint32_t fin(uint32_t v) {
    if constexpr (MAGIC) {
      uint32_t d;
      asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(v | 0x64006400u), "r"(0x64006400u));
      return d;
    } else {
      return v;
    }
  }
  static __device__ __forceinline__ void run(uint32_t x, int quad8, uint32_t (&o)[4]) {
    if constexpr (BITS == 4) {
      const uint32_t y = x >> 8;
      o[0] = fin(x & 0x000F000Fu);
      o[1] = fin(x & 0x00F000F0u);
      o[2] = fin(y & 0x000F000Fu);
      o[3] = fin(y & 0x00F000F0u);
    } else {
      const uint32_t y = x >> quad8;
      o[0] = fin(y & 0x00030003u);
      o[1] = fin(y & 0x000C000Cu);
      o[2] = fin(y & 0x00300030u);
      o[3] = fin(y & 0x00C000C0u);
    }
  }
  // 1 / (power of two the code came out with), times 2^24 when the registers were read as denormals
  static __device__ __forceinline__ float inv_factor(int i) {
    const float base = MAGIC ? 1.f : 16777216.f;
    if constexpr (BITS == 4) return (i & 1) ? base * 0.0625f : base;
    else return base * (i == 0 ? 1.f : i == 1 ? 0.25f : i == 2 ? 0.0625f : 0.015625f);
  }
};

constexpr int kv_row_stride_words(int words) { return words == 4 ? 4 : words + 4; }  // see the bank maps in the kernel

// TMA = true: tiles are fetched by cp.async.bulk.tensor with the hardware swizzle of their row width (rows unpadded, every
// LDS pattern of the contractions stays conflict-free under the XOR); false: 16 / 8-byte cp.async into padded rows (the
// fallback for 2-bit caches whose K rows are not multiples of 16 bytes, which TMA cannot address).
template <int D, int BITS, bool TMA = false>
struct KvSmem {
  static constexpr int KRB = kKvTile * BITS / 8;           // bytes of one channel row of a K tile (128 / 64)
  static constexpr int KST = TMA ? KRB / 4 : kv_row_stride_words(KRB / 4); // its stride in words
  static constexpr int VB = D * BITS / 8;                  // bytes of one token row of V (64 / 32 / 16)
  static constexpr int VST = TMA ? VB / 4 : kv_row_stride_words(VB / 4);
  static constexpr int KG = kKvTile / kKvGroup;            // K scale groups per tile (4)
  static constexpr int VG = D / kKvGroup;                  // V scale groups per token (4 / 2)
  static constexpr int kStageK = (D * KST * 4 + 1023) / 1024 * 1024, kStageV = (kKvTile * VST * 4 + 1023) / 1024 * 1024;
  static constexpr int kStage = kStageK + kStageV;
  static constexpr int kScales = (2 * KG * D + 2 * VG * kKvTile) * 2;   // one buffer: K sc, K mn, V sc, V mn (fp16, transposed)
  static constexpr int kP = kKvWarps * kKvRows * kKvPStride * 2;
  static constexpr int kEpi = kKvWarps * kKvRows * (D + 4 + 2) * 4;               // per-warp (m, l, o) at the end (reuses the ring)
  static constexpr int kRing = kKvStages * kStage > kEpi ? kKvStages * kStage : kEpi;
  static constexpr int kBytes = kRing + 2 * kScales + kP + 64 /*mbarriers*/ + 1024 /*alignment of the ring*/;
};

// Workspace layout per (b, h, row, split): [m (base 2), l, o[D]] fp32.
template <int D, int BITS, bool MAGIC, bool TMA>
__global__ void __launch_bounds__(kKvThreads, kKvCtasPerSm)
kv_attn_partial_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const __half* __restrict__ q, const uint8_t* __restrict__ kcode,
                       const __half* __restrict__ kscale, const __half* __restrict__ kmn,
                       const uint8_t* __restrict__ vcode, const __half* __restrict__ vscale,
                       const __half* __restrict__ vmn, float* __restrict__ ws, int H, int Nq, int N, int nsplit,
                       int keys_per_split, float scale_log2e, int64_t qsb, int64_t qsn, int64_t qsh, int kchunk, int dbg) {
  using SM = KvSmem<D, BITS, TMA>;
  using Dec = KvDec<BITS, MAGIC>;
  constexpr int KST = SM::KST, VST = SM::VST, KG = SM::KG, VG = SM::VG, VB = SM::VB, KRB = SM::KRB;
  constexpr int KSTEPS = D / 16;
  constexpr uint32_t kOnes = 0x3C003C00u;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  __half* scl = reinterpret_cast<__half*>(smem + SM::kRing);          // [2][K sc | K mn | V sc | V mn]
  __half* sP = reinterpret_cast<__half*>(smem + SM::kRing + 2 * SM::kScales);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kRing + 2 * SM::kScales + SM::kP);   // [kKvStages], TMA only
  const uint32_t ring_a = (uint32_t)__cvta_generic_to_shared(ring);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  // heads vary fastest across CTAs: the 32-byte sectors of V scales (4 heads) and the 2 KB of V codes of one token (all
  // heads) are then touched by co-resident CTAs within a short window (L2 sector reuse, DRAM page hits)
  const int split = blockIdx.y % nsplit, qtile = blockIdx.y / nsplit;
  const int h = blockIdx.x, b = blockIdx.z;
  const int row0 = qtile * kKvRows;
  const int n_begin = split * keys_per_split, n_end = min(N, n_begin + keys_per_split);
  const int ntiles = (n_end - n_begin + kKvTile - 1) / kKvTile;

  // lane <-> codes: which packed word of a 32-code group, which 16-bit half of it, (2-bit) which quad of that half
  const int wsel = (BITS == 4) ? (gq & 3) : (gq & 1);
  const int hsel = (BITS == 4) ? (gq >> 2) : ((gq >> 1) & 1);
  const int quad8 = (BITS == 4) ? 0 : 8 * (gq >> 2);
  const int cb = (BITS == 4) ? 8 * wsel + 4 * hsel : 16 * wsel + 8 * hsel + 4 * (gq >> 2);  // first of my 4 codes in the group
  const uint32_t prmt_sel = hsel ? 0x7632u : 0x5410u;
  constexpr int WPG = 32 * BITS / 32;  // words per 32-code group (4 / 2)

  // ---- cache addressing: kcode [B][D][H][N*BITS/8], kscale/kmn [B][D][H][N/G]; vcode [B][N][H][VB], vscale/vmn [B][N][H][VG]
  const int64_t krow_bytes = (int64_t)N * BITS / 8, kgroups = N / kKvGroup;
  const uint8_t* kc_base = kcode + ((int64_t)b * D * H + h) * krow_bytes;
  const int64_t kc_stride = (int64_t)H * krow_bytes;
  const __half* ks_base = kscale + ((int64_t)b * D * H + h) * kgroups;
  const __half* km_base = kmn + ((int64_t)b * D * H + h) * kgroups;
  const int64_t ks_stride = (int64_t)H * kgroups;
  const uint8_t* vc_base = vcode + ((int64_t)b * N * H + h) * VB;
  const int64_t vc_stride = (int64_t)H * VB;
  const __half* vs_base = vscale + ((int64_t)b * N * H + h) * VG;
  const __half* vm_base = vmn + ((int64_t)b * N * H + h) * VG;
  const int64_t vs_stride = (int64_t)H * VG;
  static_assert(KG == 8, "the K scale registers are written for 8 groups per tile");
  const bool ks_fast = ((N & 255) == 0) && ((reinterpret_cast<uintptr_t>(kscale) | reinterpret_cast<uintptr_t>(kmn)) & 15) == 0;

  // ---- staging ----
  // cp.async thread map.  A 16-byte-per-lane shared store is processed a quarter-warp at a time, so the 8 lanes of a
  // quarter take 8 CONSECUTIVE ROWS of ONE chunk column: with the padded row strides above those 8 x 16 bytes fall
  // into 32 different banks (the first version gave a quarter 2 rows x 4 columns: 27 wavefronts per instruction
  // instead of 4, and the copies alone kept the shared-memory pipe busier than all the LDS of the contractions).
  // From tile to tile a thread's source pointers advance by a constant (tiles are issued in order).
  // K: a warp-instruction covers 8 rows x 4 chunk columns; rows hold kcpr = 8 or 4 chunks, i.e. kcb = 2 or 1 column blocks
  const int kcpr = KRB / kchunk, kcb = kcpr / 4;
  const int k_col = 4 * (warp % kcb) + (lane >> 3), k_row0 = 8 * (warp / kcb) + (lane & 7);
  const uint8_t* k_src = kc_base + k_row0 * kc_stride + (int64_t)n_begin * BITS / 8 + k_col * kchunk;
  const int64_t k_step = (int64_t)(8 * kKvWarps / kcb) * kc_stride;
  int64_t k_left = krow_bytes - ((int64_t)n_begin * BITS / 8 + k_col * kchunk);   // bytes of my column before the row ends
  const uint32_t k_dst0 = (uint32_t)(k_row0 * (KST * 4) + k_col * kchunk);
  constexpr int CPV = VB / 16, VRPP = kKvThreads / CPV;  // V chunks per row, rows per pass
  const int v_col = (lane >> 3) % CPV, v_row0 = (32 / CPV) * warp + (lane & 7) + 8 * ((lane >> 3) / CPV);
  const uint8_t* v_src = vc_base + (int64_t)(n_begin + v_row0) * vc_stride + v_col * 16;
  const int64_t v_step = (int64_t)VRPP * vc_stride;
  int v_n = n_begin + v_row0;                            // token of my first row of the next tile to issue
  const uint32_t v_dst0 = (uint32_t)(v_row0 * (VST * 4) + v_col * 16);
  auto issue_codes = [&](int tile) {  // copies of the NEXT tile (call order = tile order) into ring stage tile % kKvStages
    const uint32_t sk = ring_a + (tile % kKvStages) * SM::kStage, sv = sk + SM::kStageK;
    if constexpr (TMA) {  // one thread, two boxes: K [D rows x KRB bytes] of (b, h), V [tile tokens x VB bytes]; out-of-range bytes arrive as zeros
      if (tid == 0) {
        const uint32_t bar = ptx::smem_u32(&full[tile % kKvStages]);
        const int n0 = n_begin + tile * kKvTile;
        ptx::mbar_expect_tx_a(bar, D * KRB + kKvTile * VB);
        ptx::tma_load_4d_a(sk, &tmK, bar, n0 * BITS / 8, 0, h, b);
        ptx::tma_load_4d_a(sv, &tmV, bar, 0, n0, h, b);
      }
      return;
    }
    auto k_rows = [&](auto chunk_tag) {
      constexpr int CH = decltype(chunk_tag)::value;
      constexpr int RPP = 8 * kKvWarps / (KRB / CH / 4), PASSES = D / RPP;
      static_assert(KRB / CH == 4 || KRB / CH == 8, "K rows are 4 or 8 chunks");
      uint32_t dst = sk + k_dst0;
      if (k_left >= CH) {                                // the common case: no zero fill
        const uint8_t* src = k_src;
#pragma unroll
        for (int j = 0; j < PASSES; ++j, src += k_step, dst += RPP * KST * 4) kv_cp_async_full<CH>(dst, src);
      } else {
        const int valid = k_left > 0 ? (int)k_left : 0;
        const uint8_t* src = valid > 0 ? k_src : kc_base;
        const int64_t step = valid > 0 ? k_step : 0;
#pragma unroll
        for (int j = 0; j < PASSES; ++j, src += step, dst += RPP * KST * 4) kv_cp_async<CH>(dst, src, valid);
      }
    };
    if (dbg & 2) {
    } else if constexpr (KRB / 8 > 8) {   // 4-bit rows are always multiples of 16 bytes
      k_rows(std::integral_constant<int, 16>{});
    } else {
      if (kchunk == 16) k_rows(std::integral_constant<int, 16>{});
      else k_rows(std::integral_constant<int, 8>{});
    }
    k_src += KRB;
    k_left -= KRB;
    if (!(dbg & 4)) {
      const uint8_t* src = v_src;
      uint32_t dst = sv + v_dst0;
      if (v_n + (CPV - 1) * VRPP < N) {
#pragma unroll
        for (int j = 0; j < CPV; ++j, src += v_step, dst += VRPP * VST * 4) kv_cp_async_full<16>(dst, src);
      } else {
#pragma unroll
        for (int j = 0; j < CPV; ++j, src += v_step, dst += VRPP * VST * 4) {
          const bool ok = v_n + j * VRPP < N;
          kv_cp_async<16>(dst, ok ? src : vc_base, ok ? 16 : 0);
        }
      }
    }
    v_src += (int64_t)kKvTile * vc_stride;
    v_n += kKvTile;
  };
  // scales travel through registers (global [d][group] / [token][group] -> shared [group][d] / [group][token])
  uint32_t rks[KG / 2], rkm[KG / 2];   // KG fp16 values each
  uint2 rvs = make_uint2(0, 0), rvm = make_uint2(0, 0);
#pragma unroll
  for (int g = 0; g < KG / 2; ++g) rks[g] = rkm[g] = 0u;
  auto fetch_scales = [&](int tile) {
    if (dbg & 8) return;
    const int n0 = n_begin + tile * kKvTile;
    if (tid < D) {
      const int64_t g0 = n0 / kKvGroup;
      const __half* sp = ks_base + tid * ks_stride + g0;
      const __half* mp = km_base + tid * ks_stride + g0;
      if (ks_fast) {
        const uint4 a = *reinterpret_cast<const uint4*>(sp), m4 = *reinterpret_cast<const uint4*>(mp);
        rks[0] = a.x, rks[1] = a.y, rks[2] = a.z, rks[3] = a.w;
        rkm[0] = m4.x, rkm[1] = m4.y, rkm[2] = m4.z, rkm[3] = m4.w;
      } else {
#pragma unroll
        for (int g = 0; g < KG / 2; ++g) {
          const bool ok0 = g0 + 2 * g < kgroups, ok1 = g0 + 2 * g + 1 < kgroups;
          const uint32_t a0 = ok0 ? *reinterpret_cast<const uint16_t*>(sp + 2 * g) : 0u;
          const uint32_t a1 = ok1 ? *reinterpret_cast<const uint16_t*>(sp + 2 * g + 1) : 0u;
          const uint32_t m0 = ok0 ? *reinterpret_cast<const uint16_t*>(mp + 2 * g) : 0u;
          const uint32_t m1 = ok1 ? *reinterpret_cast<const uint16_t*>(mp + 2 * g + 1) : 0u;
          rks[g] = a0 | (a1 << 16);
          rkm[g] = m0 | (m1 << 16);
        }
      }
    }
    {
      const int n = n0 + tid;
      const bool ok = n < N;
      const __half* sp = vs_base + (ok ? n : 0) * vs_stride;
      const __half* mp = vm_base + (ok ? n : 0) * vs_stride;
      if constexpr (VG == 4) {
        rvs = ok ? *reinterpret_cast<const uint2*>(sp) : make_uint2(0, 0);
        rvm = ok ? *reinterpret_cast<const uint2*>(mp) : make_uint2(0, 0);
      } else {
        rvs.x = ok ? *reinterpret_cast<const uint32_t*>(sp) : 0u;
        rvm.x = ok ? *reinterpret_cast<const uint32_t*>(mp) : 0u;
      }
    }
  };
  // Shared layout of one scale buffer: 16-byte units {sc(k), sc(k+1), sc(k+8), sc(k+9), mn(k), mn(k+1), mn(k+8), mn(k+9)} --
  // exactly the fp16 pairs that multiply one lane's two B registers of one MMA k-step -- so a k-step costs ONE LDS.128.
  // K: unit (group g, k-step ks, t) at g * D/4 + 4 ks + t, k = channel 16 ks + 2 t;  V: unit (group g, 16-token block nb, t)
  // at KG * D/4 + g * (tile / 4) + 4 nb + t, k = token 16 nb + 2 t of the tile.
  auto unit_slot = [](int e) { return 2 * ((e >> 3) & 1) + (e & 1); };   // position of element e (mod 16) inside its unit
  auto park_scales = [&](int buf) {
    uint16_t* base = reinterpret_cast<uint16_t*>(scl) + buf * (SM::kScales / 2);
    if (tid < D) {
      uint16_t* u = base + ((tid >> 4) * 4 + ((tid & 7) >> 1)) * 8 + unit_slot(tid);
#pragma unroll
      for (int g = 0; g < KG; ++g) {
        u[g * (D / 4) * 8] = (uint16_t)((g & 1) ? rks[g / 2] >> 16 : rks[g / 2] & 0xffff);
        u[g * (D / 4) * 8 + 4] = (uint16_t)((g & 1) ? rkm[g / 2] >> 16 : rkm[g / 2] & 0xffff);
      }
    }
    {
      uint16_t* u = base + KG * (D / 4) * 8 + ((tid >> 4) * 4 + ((tid & 7) >> 1)) * 8 + unit_slot(tid);
      const uint32_t sc[4] = {rvs.x & 0xffff, rvs.x >> 16, rvs.y & 0xffff, rvs.y >> 16};
      const uint32_t mn[4] = {rvm.x & 0xffff, rvm.x >> 16, rvm.y & 0xffff, rvm.y >> 16};
#pragma unroll
      for (int g = 0; g < VG; ++g) {
        u[g * (kKvTile / 4) * 8] = (uint16_t)sc[g];
        u[g * (kKvTile / 4) * 8 + 4] = (uint16_t)mn[g];
      }
    }
  };

  if constexpr (TMA) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < kKvStages; ++s) ptx::mbar_init(&full[s], 1);
      ptx::fence_barrier_init();
      ptx::prefetch_tmap(&tmK);
      ptx::prefetch_tmap(&tmV);
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < kKvStages - 1; ++s) {
    if (s < ntiles) issue_codes(s);
    kv_cp_commit();
  }
  fetch_scales(0);

  // ---- query fragment: B operand rows (k = channel pair) of query row gq, pre-multiplied by softmax_scale * log2(e)
  uint32_t q2[KSTEPS][2];
  {
    const bool row_ok = row0 + gq < Nq;
    const __half* qp = q + b * qsb + (int64_t)(row_ok ? row0 + gq : 0) * qsn + h * qsh;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int d0 = 16 * ks + 8 * hh + 2 * t;
        const float a = row_ok ? __half2float(qp[d0]) * scale_log2e : 0.f;
        const float c = row_ok ? __half2float(qp[d0 + 1]) * scale_log2e : 0.f;
        q2[ks][hh] = kv_pack_f16x2(a, c);
      }
    }
  }
  park_scales(0);
  if (ntiles > 1) fetch_scales(1);

  // running state of this warp's key slice: rows 2t, 2t+1 (the accumulator columns this lane owns)
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float accO[VG][2][4], accV[VG][4];
#pragma unroll
  for (int g = 0; g < VG; ++g) {
#pragma unroll
    for (int c = 0; c < 4; ++c) accO[g][0][c] = accO[g][1][c] = accV[g][c] = 0.f;
  }
  __half* sPw = sP + warp * (kKvRows * kKvPStride);
  // per-lane byte offsets inside a stage / a scale buffer / the P tile; everything else is an immediate
  const uint32_t scl_a0 = (uint32_t)__cvta_generic_to_shared(scl);
  // byte offset of my word in row `row` of a tile whose rows are RB bytes: plain (padded rows), or under the TMA
  // swizzle of that row width -- address bits [4, 4 + log2(RB/16)) ^= bits [7, ...)
  auto word_off = [](int row, int cw, int rb, int stride_words) -> uint32_t {
    if constexpr (!TMA) return (uint32_t)(row * stride_words + cw) * 4u;
    const int f = ((row * rb) >> 7) & (rb / 16 - 1);
    return (uint32_t)(row * rb + ((((cw >> 2) ^ f)) << 4) + (cw & 3) * 4);
  };
  const int kcw = warp * WPG + wsel;
  // rows 2t, 2t+8 (+16 ks) share one swizzle phase, rows 2t+1, 2t+9 another (they differ only for 128-byte rows)
  const uint32_t k_lane_off_e = word_off(2 * t, kcw, KRB, KST);
  const uint32_t k_lane_off_o = word_off(2 * t + 1, kcw, KRB, KST) - (uint32_t)(KST * 4);
  uint32_t v_lane_off[VG];   // per channel group; the four token rows of a k-step share one swizzle phase (VB <= 64)
#pragma unroll
  for (int g = 0; g < VG; ++g) v_lane_off[g] = word_off(warp * 32 + 2 * t, g * WPG + wsel, VB, VST);
  const uint32_t ks_lane_off = (uint32_t)(warp * (D / 4) + t) * 16u;
  const uint32_t vs_lane_off = (uint32_t)(KG * (D / 4) + 8 * warp + t) * 16u;
  const uint32_t p_rd = (uint32_t)__cvta_generic_to_shared(sPw) + (uint32_t)(gq * kKvPStride + 2 * t) * 2u;

  for (int i = 0; i < ntiles; ++i) {
    if constexpr (TMA) ptx::mbar_wait(&full[i % kKvStages], (uint32_t)(i / kKvStages) & 1u, 40);
    else kv_cp_wait<kKvStages - 2>();
    __syncthreads();  // tile i landed for everyone; tile i-1 fully consumed; scale buffer i&1 visible
    if (i + kKvStages - 1 < ntiles) issue_codes(i + kKvStages - 1);
    kv_cp_commit();
    if (i + 1 < ntiles) park_scales((i + 1) & 1);
    if (i + 2 < ntiles) fetch_scales(i + 2);
    if (dbg & 1) continue;   // LOWBIT_KV_DEBUG timing experiments (results are garbage)

    const int n0 = n_begin + i * kKvTile;
    const uint32_t stage_a = ring_a + (i % kKvStages) * SM::kStage;
    const uint32_t scl_a = scl_a0 + (i & 1) * SM::kScales;

    // ---- scores of my 32 keys: S^T (keys x rows) = codes^T . (q * sc), + ones . (q * mn) ----
    float accS[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, accC[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint32_t ka = stage_a + k_lane_off_e, kb = stage_a + k_lane_off_o;   // my word of rows 2t (+8) / 2t+1 (+9)
      const uint32_t sa = scl_a + ks_lane_off;             // unit (group = warp, k-step 0, t)
      struct KLoad { uint4 sm; uint32_t w00, w01, w10, w11; };
      auto kload = [&](auto ks_tag) {
        constexpr int ks = decltype(ks_tag)::value;
        constexpr int RO = 16 * ks * KST * 4;
        KLoad L;
        L.sm = kv_lds128<ks * 64>(sa);
        L.w00 = kv_lds32<RO>(ka), L.w01 = kv_lds32<RO + KST * 4>(kb);
        L.w10 = kv_lds32<RO + 8 * KST * 4>(ka), L.w11 = kv_lds32<RO + 9 * KST * 4>(kb);
        return L;
      };
      auto kmath = [&](auto ks_tag, const KLoad& L) {
        constexpr int ks = decltype(ks_tag)::value;
        uint32_t lo[4], hi[4];
        Dec::run(__byte_perm(L.w00, L.w01, prmt_sel), quad8, lo);
        Dec::run(__byte_perm(L.w10, L.w11, prmt_sel), quad8, hi);
        const uint32_t b0 = kv_hmul2(q2[ks][0], L.sm.x), b1 = kv_hmul2(q2[ks][1], L.sm.y);
        kv_mma(accS[0], lo[0], lo[1], hi[0], hi[1], b0, b1);
        kv_mma(accS[1], lo[2], lo[3], hi[2], hi[3], b0, b1);
        kv_mma(accC, kOnes, kOnes, kOnes, kOnes, kv_hmul2(q2[ks][0], L.sm.z), kv_hmul2(q2[ks][1], L.sm.w));
      };
      // loads run one k-step ahead of the arithmetic
#define KV_KSTEP(A, B, LA, LB) KLoad LB = kload(std::integral_constant<int, B>{}); kmath(std::integral_constant<int, A>{}, LA);
      KLoad L0 = kload(std::integral_constant<int, 0>{});
      KV_KSTEP(0, 1, L0, L1)
      KV_KSTEP(1, 2, L1, L2)
      KV_KSTEP(2, 3, L2, L3)
      if constexpr (KSTEPS == 8) {
        KV_KSTEP(3, 4, L3, L4)
        KV_KSTEP(4, 5, L4, L5)
        KV_KSTEP(5, 6, L5, L6)
        KV_KSTEP(6, 7, L6, L7)
        kmath(std::integral_constant<int, 7>{}, L7);
      } else {
        kmath(std::integral_constant<int, 3>{}, L3);
      }
#undef KV_KSTEP
    }
    // ---- online softmax (base 2), rows 2t + j; the 32 keys of a row sit in the 8 lanes that share t ----
    float p[4][2], alpha[2];
    const int key0 = n0 + warp * 32 + cb;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float s[4];
      s[0] = fmaf(accS[0][j], Dec::inv_factor(0), accC[j]);
      s[1] = fmaf(accS[0][2 + j], Dec::inv_factor(1), accC[j]);
      s[2] = fmaf(accS[1][j], Dec::inv_factor(2), accC[j]);
      s[3] = fmaf(accS[1][2 + j], Dec::inv_factor(3), accC[j]);
      float mx = -INFINITY;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        s[k4] = (key0 + k4 < N) ? s[k4] : -INFINITY;
        mx = fmaxf(mx, s[k4]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      const float m_new = fmaxf(m_run[j], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // a slice with no live key yet
      alpha[j] = kv_ex2(m_run[j] - m_use);
      m_run[j] = m_new;
      float lsum = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        p[k4][j] = kv_ex2(s[k4] - m_use);
        lsum += p[k4][j];
      }
      l_run[j] = fmaf(l_run[j], alpha[j], lsum);
      // P^T tile of this warp: [row][key] fp16; my 4 keys of row 2t+j are 8 contiguous bytes
      *reinterpret_cast<uint2*>(sPw + (2 * t + j) * kKvPStride + cb) =
          make_uint2(kv_pack_f16x2(p[0][j], p[1][j]), kv_pack_f16x2(p[2][j], p[3][j]));
    }
    if (__any_sync(0xffffffffu, alpha[0] != 1.f || alpha[1] != 1.f)) {
#pragma unroll
      for (int g = 0; g < VG; ++g) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          accO[g][0][c] *= alpha[c & 1];
          accO[g][1][c] *= alpha[c & 1];
          accV[g][c] *= alpha[c & 1];
        }
      }
    }
    __syncwarp();
    // ---- P.V over my 32 keys: O^T (channels x rows) += codes^T . (p * vs), + ones . (p * vm) ----
    {
      const uint32_t vbase = stage_a + SM::kStageK;             // + v_lane_off[g]: word (token 32 w + 2 t, my column of group g)
      const uint32_t sa = scl_a + vs_lane_off;                  // unit (group 0, token block 2 w, t)
      struct VLoad { uint4 sm; uint32_t w00, w01, w10, w11; };
      auto vload = [&](auto k2_tag, auto g_tag) {
        constexpr int k2 = decltype(k2_tag)::value, g = decltype(g_tag)::value;
        constexpr int RO = 16 * k2 * VST * 4;
        const uint32_t va = vbase + v_lane_off[g];
        VLoad L;
        L.sm = kv_lds128<(g * (kKvTile / 4) + 4 * k2) * 16>(sa);
        L.w00 = kv_lds32<RO>(va), L.w01 = kv_lds32<RO + VST * 4>(va);
        L.w10 = kv_lds32<RO + 8 * VST * 4>(va), L.w11 = kv_lds32<RO + 9 * VST * 4>(va);
        return L;
      };
      auto vmath = [&](auto g_tag, const VLoad& L, uint32_t p_a, uint32_t p_b) {
        constexpr int g = decltype(g_tag)::value;
        uint32_t lo[4], hi[4];
        Dec::run(__byte_perm(L.w00, L.w01, prmt_sel), quad8, lo);
        Dec::run(__byte_perm(L.w10, L.w11, prmt_sel), quad8, hi);
        const uint32_t b0 = kv_hmul2(p_a, L.sm.x), b1 = kv_hmul2(p_b, L.sm.y);
        kv_mma(accO[g][0], lo[0], lo[1], hi[0], hi[1], b0, b1);
        kv_mma(accO[g][1], lo[2], lo[3], hi[2], hi[3], b0, b1);
        kv_mma(accV[g], kOnes, kOnes, kOnes, kOnes, kv_hmul2(p_a, L.sm.z), kv_hmul2(p_b, L.sm.w));
      };
      using I0 = std::integral_constant<int, 0>;
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      using I3 = std::integral_constant<int, 3>;
      const uint32_t pa0 = kv_lds32<0>(p_rd), pb0 = kv_lds32<16>(p_rd), pa1 = kv_lds32<32>(p_rd), pb1 = kv_lds32<48>(p_rd);
      VLoad A = vload(I0{}, I0{});
      VLoad B = vload(I0{}, I1{});
      vmath(I0{}, A, pa0, pb0);
      if constexpr (VG == 4) {
        A = vload(I0{}, I2{});
        vmath(I1{}, B, pa0, pb0);
        B = vload(I0{}, I3{});
        vmath(I2{}, A, pa0, pb0);
        A = vload(I1{}, I0{});
        vmath(I3{}, B, pa0, pb0);
        B = vload(I1{}, I1{});
        vmath(I0{}, A, pa1, pb1);
        A = vload(I1{}, I2{});
        vmath(I1{}, B, pa1, pb1);
        B = vload(I1{}, I3{});
        vmath(I2{}, A, pa1, pb1);
        vmath(I3{}, B, pa1, pb1);
      } else {
        A = vload(I1{}, I0{});
        vmath(I1{}, B, pa0, pb0);
        B = vload(I1{}, I1{});
        vmath(I0{}, A, pa1, pb1);
        vmath(I1{}, B, pa1, pb1);
      }
    }
  }
  kv_cp_wait<0>();

  // ---- the four warps' states meet once: (m, l, o) of the split ----
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    l_run[j] += __shfl_xor_sync(0xffffffffu, l_run[j], 4);
    l_run[j] += __shfl_xor_sync(0xffffffffu, l_run[j], 8);
    l_run[j] += __shfl_xor_sync(0xffffffffu, l_run[j], 16);
  }
  __syncthreads();  // the ring is free
  float* eM = reinterpret_cast<float*>(ring);          // [warps][8]
  float* eL = eM + kKvWarps * kKvRows;                 // [warps][8]
  float* eO = eL + kKvWarps * kKvRows;                 // [warps][8][D + 4]
  if (gq == 0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      eM[warp * kKvRows + 2 * t + j] = m_run[j];
      eL[warp * kKvRows + 2 * t + j] = l_run[j];
    }
  }
#pragma unroll
  for (int g = 0; g < VG; ++g) {
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int code = 2 * mm + (c >> 1);            // which of my 4 channels
        const int ch = g * 32 + cb + code, r = 2 * t + (c & 1);
        eO[(warp * kKvRows + r) * (D + 4) + ch] = fmaf(accO[g][mm][c], Dec::inv_factor(code), accV[g][c & 1]);
      }
    }
  }
  __syncthreads();
  const int rows = min(kKvRows, Nq - row0);
  for (int idx = tid; idx < rows * D; idx += kKvThreads) {
    const int r = idx / D, c = idx % D;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kKvWarps; ++w) M = fmaxf(M, eM[w * kKvRows + r]);
    const float Mu = (M == -INFINITY) ? 0.f : M;
    float acc = 0.f, lsum = 0.f;
#pragma unroll
    for (int w = 0; w < kKvWarps; ++w) {
      const float wt = kv_ex2(eM[w * kKvRows + r] - Mu);
      acc = fmaf(eO[(w * kKvRows + r) * (D + 4) + c], wt, acc);
      lsum = fmaf(eL[w * kKvRows + r], wt, lsum);
    }
    float* dst = ws + ((((int64_t)b * H + h) * Nq + (row0 + r)) * nsplit + split) * (D + 2);
    dst[2 + c] = acc;
    if (c == 0) {
      dst[0] = M;
      dst[1] = lsum;
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

static int kv_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static int kv_splits(int B, int H, int Nq, int N, int* keys_per_split) {
  // splits of whole tiles, none of them empty.  LOWBIT_KV_SPLITS forces a count (tuning aid).
  const int tiles = (N + kKvTile - 1) / kKvTile;
  const int64_t base = (int64_t)B * H * ((Nq + kKvRows - 1) / kKvRows);
  static const int forced = kv_env_int("LOWBIT_KV_SPLITS", 0);
  int want = forced > 0 ? forced : (int)((148 * kKvCtasPerSm * 4) / base);
  want = want < 1 ? 1 : (want > tiles ? tiles : want);
  const int tps = (tiles + want - 1) / want;  // tiles per split
  *keys_per_split = tps * kKvTile;
  return (tiles + tps - 1) / tps;
}

template <int D, int BITS, bool MAGIC, bool TMA>
static int kv_launch(dim3 grid, cudaStream_t st, const void* q, const void* kcode, const void* kscale, const void* kmn,
                     const void* vcode, const void* vscale, const void* vmn, void* workspace, int B, int H, int Nq, int N,
                     int ns, int kps, float sl2, int64_t qsb, int64_t qsn, int64_t qsh, int kchunk) {
  const int dbg = kv_env_int("LOWBIT_KV_DEBUG", 0);  // timing experiments only: 1 no arithmetic, 2 no K codes, 4 no V codes, 8 no scales
  auto kern = kv_attn_partial_kernel<D, BITS, MAGIC, TMA>;
  using SM = KvSmem<D, BITS, TMA>;
  constexpr int bytes = SM::kBytes;
  CUtensorMap tmK, tmV;
  memset(&tmK, 0, sizeof(tmK));
  memset(&tmV, 0, sizeof(tmV));
  if (TMA) {
    auto swz = [](int row_bytes) {
      return row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    };
    const int64_t krow = (int64_t)N * BITS / 8;
    // K codes [B][D][H][krow bytes] as (byte, d, h, b); V codes [B][N][H][VB bytes] as (byte, n, h, b)
    const int64_t kdim[4] = {krow, D, H, B}, kstr[3] = {(int64_t)H * krow, krow, (int64_t)D * H * krow};
    if (make_map(&tmK, kcode, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, kdim, kstr, SM::KRB, D, swz(SM::KRB)) != 0) return 1;
    const int64_t vdim[4] = {SM::VB, N, H, B}, vstr[3] = {(int64_t)H * SM::VB, SM::VB, (int64_t)N * H * SM::VB};
    if (make_map(&tmV, vcode, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, vdim, vstr, SM::VB, kKvTile, swz(SM::VB)) != 0) return 1;
  }
  // the opt-in is per device: set before every launch (cheap), not behind a process-wide flag
  LOWBIT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  kern<<<grid, kKvThreads, bytes, st>>>(tmK, tmV, (const __half*)q, (const uint8_t*)kcode, (const __half*)kscale,
                                        (const __half*)kmn, (const uint8_t*)vcode, (const __half*)vscale, (const __half*)vmn,
                                        (float*)workspace, H, Nq, N, ns, kps, sl2, qsb, qsn, qsh, kchunk, dbg);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace lowbit

using namespace lowbit;

extern "C" int64_t lowbit_kv_attn_workspace_bytes(int B, int H, int Nq, int N, int D) {
  int kps = 0;
  const int ns = kv_splits(B, H, Nq, N, &kps);
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
  LOWBIT_CHECK(((uintptr_t)kcode & 15) == 0, "lowbit_kv_attn_fwd: K codes must be 16-byte aligned");
  LOWBIT_CHECK(((uintptr_t)vcode & 15) == 0, "lowbit_kv_attn_fwd: V codes must be 16-byte aligned");
  LOWBIT_CHECK((((uintptr_t)vscale | (uintptr_t)vmn) & 7) == 0, "lowbit_kv_attn_fwd: V scales / minima must be 8-byte aligned");
  LOWBIT_CHECK(lse == nullptr || lse_stride >= Nq, "lowbit_kv_attn_fwd: lse_stride < Nq");
  cudaStream_t st = (cudaStream_t)stream;
  int kps = 0;
  const int ns = kv_splits(B, H, Nq, N, &kps);
  const int qtiles = (Nq + kKvRows - 1) / kKvRows;
  dim3 grid(H, (unsigned)(ns * qtiles), B);
  const float sl2 = softmax_scale * 1.4426950408889634f;
  // widest cp.async that every K row start allows: rows are N * bits / 8 bytes apart
  const int64_t krow = (int64_t)N * bits / 8;
  const int kchunk = (krow & 15) == 0 ? 16 : 8;   // N % 32 == 0: 4-bit rows are multiples of 16 bytes, 2-bit rows of 8
  const bool magic = kv_env_int("LOWBIT_KV_MAGIC", 0) != 0;  // 1: never hand fp16 denormals to the tensor core
  // TMA needs every stride to be a multiple of 16 bytes: true for 4-bit caches (N % 32 == 0), and for 2-bit ones when
  // N % 64 == 0; the rest (and the MAGIC form, a validation aid) take the cp.async path
  const bool tma = !magic && (krow & 15) == 0 && kv_env_int("LOWBIT_KV_TMA", 1) != 0;
  int rc;
#define KV_ARGS grid, st, q, kcode, kscale, kmn, vcode, vscale, vmn, workspace, B, H, Nq, N, ns, kps, sl2, qsb, qsn, qsh, kchunk
#define KV_LAUNCH(DD, BB)                                                                                              \
  rc = tma ? kv_launch<DD, BB, false, true>(KV_ARGS)                                                                  \
           : (magic ? kv_launch<DD, BB, true, false>(KV_ARGS) : kv_launch<DD, BB, false, false>(KV_ARGS))
  if (D == 64) { if (bits == 4) KV_LAUNCH(64, 4); else KV_LAUNCH(64, 2); }
  else { if (bits == 4) KV_LAUNCH(128, 4); else KV_LAUNCH(128, 2); }
#undef KV_ARGS
#undef KV_LAUNCH
  if (rc != 0) return rc;
  dim3 mgrid(Nq, H, B);
  if (D == 64)
    kv_attn_merge_kernel<64><<<mgrid, 64, 0, st>>>((const float*)workspace, (__half*)o, lse, H, Nq, ns, osb, osn, osh, lse_stride);
  else
    kv_attn_merge_kernel<128><<<mgrid, 128, 0, st>>>((const float*)workspace, (__half*)o, lse, H, Nq, ns, osb, osn, osh, lse_stride);
  LOWBIT_CUDA(cudaGetLastError());
  return 0;
}
