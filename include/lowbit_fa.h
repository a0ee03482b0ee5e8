/* lowbit_fa.h -- C ABI of liblowbit_fa_b200.so (hand-written sm_100a CUDA).
 *
 * This is the drop-in boundary for the reference's low-bit FlashAttention hot path
 * (Charles2530/lowbit_quant_fa2_paddle; paths below are relative to that repository).
 * Every entry point takes plain device pointers, sizes, element strides and a cudaStream_t
 * (as void*).  No framework types cross this boundary; the Python host code in
 * lowbit_quant_fa2_paddle_b200/ obtains pointers from Paddle / torch tensors (DLPack hand-off)
 * and calls these through ctypes.
 *
 * Conventions
 *   - tensors are addressed as logical [B, H, N, D] with element strides (stride_b, stride_h,
 *     stride_n); the last dimension is contiguous (src/core.py:288-290).  HND and NHD layouts of
 *     the reference differ only in the strides the host passes (quant_per_block.py:188-201).
 *   - every function returns 0 on success, non-zero on error; lowbit_last_error() returns a
 *     thread-local message.  Nothing allocates device memory: outputs and scratch belong to
 *     the caller (reference: paddle.empty in the Python wrappers).
 *   - nothing synchronises the stream or the device.
 */
#ifndef LOWBIT_FA_H_
#define LOWBIT_FA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOWBIT_ABI_VERSION 8

/* element types of the floating-point inputs / outputs */
enum { LOWBIT_F16 = 0, LOWBIT_BF16 = 1 };

/* rounding / arithmetic conventions of the per-block quantizers */
enum {
  LOWBIT_QMODE_TRITON = 0, /* Q1: src/triton/quant_per_block.py:132-178 -- x/scale, +-0.5, trunc; no eps;
                                  K smoothing subtracts in the input dtype (:186-187) */
  LOWBIT_QMODE_CUDA = 1    /* Q2: csrc/fused/fused.cu:64-198 -- amax floor 1e-7, x*(127/amax), RNE+sat;
                                  K smoothing subtracts in fp32 (:120-126) */
};

/* validation switch, OR-ed into `mode`: take the Q1 quotient x/scale from the IEEE division instead of the
 * three-instruction correctly rounded sequence (reciprocal, exact FMA remainder, corrected FMA) the kernel
 * normally uses; tests assert both give identical codes. */
#define LOWBIT_QMODE_FLAG_IEEE_DIV 0x100

/* Q1 as the reference's Triton kernels behave when JIT-compiled for a GPU (not under the interpreter): fp32 `/`
 * lowers to PTX div.full.f32 (approximate, <= 2 ulp) for `scale = max|x| / 127` and for `x / scale`
 * (quant_per_block.py:173-174).  OR-ed into LOWBIT_QMODE_TRITON.  Codes differ from the IEEE ones by at most one
 * step in a small fraction of positions; verified bit-exact against the JIT-compiled reference kernels on B200
 * (tools/ref_on_b200.py, profiles/r1_reference_on_b200.jsonl). */
#define LOWBIT_QMODE_FLAG_DIV_FULL 0x200

/* Q*K^T operand formats of the attention kernel */
enum {
  LOWBIT_QK_I8 = 0,     /* Q int8, K int8 (one code per byte)                               */
  LOWBIT_QK_Q8K4 = 1,   /* Q int8, K int4 packed two codes per byte (low nibble = even d)   */
  LOWBIT_QK_Q8KMIX = 2, /* Q int8, K per-64-block bit width from `kbits` (8/4/2) in the mixed
                           container written by lowbit_quant_k_mixed                          */
  LOWBIT_QK_F16 = 3     /* no quantization: q_codes / k_codes are the fp16 / bf16 tensors themselves (in the
                           output dtype), QK^T on tcgen05 kind::f16 with fp32 scores -- the "FP16" class of
                           lowbit_fa_multi_precision (src/core.py:1075-1076, `default_attn`).  q_scale points
                           at ONE float, sm_scale * log2(e); k_scale is not read.  fp16 P.V, padded tensors. */
};
enum { LOWBIT_PV_F16 = 0, LOWBIT_PV_E4M3 = 1 };

/* attention flags */
enum {
  LOWBIT_ATTN_CAUSAL = 1,      /* attn_qk_int8_per_block_causal.py:24-79 (requires Nq == Nk)          */
  LOWBIT_ATTN_COMPAT_TAIL = 2, /* reproduce the reference's unmasked tail keys when Nk % 64 != 0
                                  (attn_qk_int8_per_block.py:48-49)                  */
  LOWBIT_ATTN_NARROW = 4       /* head_dim 64: run the 32-key-step kernel instead of the default 64-key-step one
                                  (A/B and bit-identity checks against the mixed-width K path, which always
                                  runs 32-key steps); same results within the documented tolerance        */
};

int lowbit_version(void);
const char* lowbit_last_error(void);

/* Varlen (packed [T,H,D] tensors + cu_seqlens, src/core.py:356-491) -- no host synchronisation: like the reference's
 * kernels (quant_per_block_varlen.py:120, attn_qk_int8_block_varlen.py:126-128) the grids are sized from max_seqlen and a
 * CTA past its sequence's end exits.  cu_seqlens / cu_scale: int32 [nseq+1] on the device; cu_scale[b] = number of
 * scale blocks of the sequences before b (cumsum of ceil(len/blk)).  Scales are head-major: scale[h][cu_scale[b] + j]
 * with `scale_stride` entries per head (>= cu_scale[nseq]; T/blk + nseq always suffices).  km: [H,D] (ONE mean over all
 * packed tokens, core.py:448) or NULL.  Strides in elements of (head, token). */
int lowbit_quant_per_block_varlen(const void* in, const void* km, void* codes, float* scale, const int32_t* cu_seqlens,
                                  const int32_t* cu_scale, int nseq, int H, int max_seqlen, int D, int64_t ish,
                                  int64_t isn, int64_t osh, int64_t osn, int scale_stride, int blk, int bits, int pack,
                                  float sm_scale_arg, int mode, int dtype, void* stream);
/* Attention over packed codes: sequence b attends q rows cu_seqlens_q[b].. to k/v rows cu_seqlens_k[b]..; FP16 P.V;
 * tail keys are masked (LOWBIT_ATTN_COMPAT_TAIL is rejected: the rows after a sequence belong to the next one);
 * causal needs q_len == k_len per sequence (not checked: that would need a device->host read).  A sequence without
 * keys gets zeros (attn_qk_int8_block_varlen.py:168-189).  o: packed [Tq,Hq,D]. */
int lowbit_attn_fwd_varlen(const void* q_codes, const void* k_codes, const void* v, const float* q_scale,
                           const float* k_scale, const int32_t* kbits, const int32_t* cu_seqlens_q,
                           const int32_t* cu_seqlens_k, const int32_t* cu_q_scale, const int32_t* cu_k_scale, void* o,
                           int nseq, int Hq, int Hkv, int Tq, int Tk, int max_seqlen_q, int D, int64_t qsh, int64_t qsn,
                           int64_t ksh, int64_t ksn, int64_t vsh, int64_t vsn, int64_t osh, int64_t osn,
                           int q_scale_stride, int k_scale_stride, int qk_mode, int out_dtype, int flags, void* stream);

/* Dynamic K bit allocation (the reference has thresholds, core.py:1055-1061, and harnesses but no kernel;
 * semantics stated here, parity unpinned): every 64-row block of (k - km) is quantized symmetrically to INT8, INT4 or
 * INT2 (codes in [-127,127] / [-7,7] / [-1,1], Q1 or Q2 rounding per `mode`).  The width is kbits_in[b,h,j] when given,
 * else chosen from the block statistic st = max|k - km| / 127:  st > thr8 -> 8,  st > thr4 -> 4,  else 2.
 * codes: a container of D bytes per row (strides like any int8 [.., D] tensor, 16-byte aligned); a block of width w
 * uses the first D*w/8 bytes of each of its rows, in the byte order the attention kernel expands in shared memory
 * (csrc/quant.cu, "mixed-width K").  scale, kbits_out: [B,H,ceil(N/64)] contiguous.  Consumed by lowbit_attn_fwd /
 * lowbit_attn_fwd_partial with qk_mode LOWBIT_QK_Q8KMIX, whose K loads move D*w/8 bytes per row. */
int lowbit_quant_k_mixed(const void* k, const void* km, const int32_t* kbits_in, void* codes, float* scale,
                         int32_t* kbits_out, int B, int H, int N, int D, int64_t isb, int64_t ish, int64_t isn,
                         int64_t osb, int64_t osh, int64_t osn, float thr8, float thr4, int mode, int dtype,
                         void* stream);

/* S1 -- K mean over the sequence (src/core.py:293 `k.mean(dim=seq_dim, keepdim=True)`).
 * km_out: [B, H, D] contiguous, same dtype as k.  workspace: >= lowbit_k_mean_workspace_bytes().
 * fp16: exact fixed-point sum (order independent) -> fp32 -> / N -> fp16; bf16: fp64 partial sums. */
int64_t lowbit_k_mean_workspace_bytes(int B, int H, int N, int D);
int lowbit_k_mean(const void* k, void* km_out, void* workspace, int B, int H, int N, int D,
                  int64_t stride_b, int64_t stride_h, int64_t stride_n, int dtype, void* stream);

/* S1 + Q1/Q2/Q4 for K in ONE launch and ONE pass over HBM: K mean over the sequence (src/core.py:293), `k - km` and the
 * per-64-row-block codes of the smoothed K (quant_per_block.py:181-248 with km / src/quant.py:70-98).  A thread-block
 * cluster of 8 CTAs holds one (batch, kv-head) slice in shared memory, reduces the exact column sums over distributed
 * shared memory and quantizes from there.  km_out [B,H,D] (input dtype), codes / scale as lowbit_quant_per_block with
 * blk = 64: bit-identical to lowbit_k_mean followed by lowbit_quant_per_block(k, km, blk=64, sm=1).
 * fp16 only, and a slice has to fit the cluster (N * D * 2 bytes <= ~1.6 MB): lowbit_k_smooth_quant_supported() says
 * so; otherwise call the two separate entry points. */
int lowbit_k_smooth_quant_supported(int N, int D, int dtype);
int lowbit_k_smooth_quant(const void* k, void* km_out, void* codes, float* scale, int B, int H, int N, int D,
                          int64_t isb, int64_t ish, int64_t isn, int64_t osb, int64_t osh, int64_t osn,
                          int bits, int pack, int mode, int dtype, void* stream);

/* Q1/Q2/Q4 (+INT2) -- symmetric per-block quantizer.
 * Replaces: quant_per_block_int8_kernel / quant_per_block_int4_unpack_kernel
 *           (src/triton/quant_per_block.py:132-178, :22-71) and QuantInt8Kernel
 *           (csrc/fused/fused.cu:64-198; launchers :430-683, pybind csrc/fused/pybind.cpp:21-33).
 * in: [B,H,N,D] fp16/bf16; km: NULL or [B,H,D] (same dtype) subtracted before quantizing (K smoothing);
 * codes: int8, same logical shape with its own strides; if `pack` != 0 and bits < 8 the last dim holds
 * D*bits/8 bytes (strides are then in bytes of the packed tensor); scale: f32 [B,H,ceil(N/blk)] contiguous.
 * x = f32(in) [- km] * sm_scale_arg;  bits in {8,4,2} -> QMAX {127,7,1}. D in {64,128}; blk in {32,64,128}. */
int lowbit_quant_per_block(const void* in, const void* km, void* codes, float* scale,
                           int B, int H, int N, int D,
                           int64_t in_stride_b, int64_t in_stride_h, int64_t in_stride_n,
                           int64_t out_stride_b, int64_t out_stride_h, int64_t out_stride_n,
                           int blk, int bits, int pack, float sm_scale_arg, int mode, int dtype, void* stream);

/* Q3 -- per-thread-group quantizer (src/triton/quant_per_thread.py:22-219, hosts :222-411).
 * is_key == 0: groups of rows {8i+t} inside each `warp_blk`(32)-row block, 8 scales per block;
 * is_key == 1: groups of rows {8i+2t, 8i+2t+1} inside each `warp_blk`(64)-row block, 4 scales per block.
 * scale = amax/QMAX + 1e-7; n_scale = number of scale slots per (b,h) (slots past the data get 1e-7). */
int lowbit_quant_per_thread(const void* in, const void* km, void* codes, float* scale,
                            int B, int H, int N, int D,
                            int64_t in_stride_b, int64_t in_stride_h, int64_t in_stride_n,
                            int64_t out_stride_b, int64_t out_stride_h, int64_t out_stride_n,
                            int warp_blk, int n_scale, int is_key, int bits, int dtype, void* stream);

/* Q5 -- KIVI asymmetric group quantize + pack along the last dim
 * (src/triton/utils/quant/new_pack.py:198-300).  data: [rows, T] fp16 contiguous; group = 32;
 * bits in {2,4,8}; code: int8 [rows, T*bits/8]; scale, mn: fp16 [rows, T/group]. */
int lowbit_quant_pack_lastdim(const void* data, void* code, void* scale, void* mn,
                              int64_t rows, int T, int group, int bits, int dtype, void* stream);

/* KV-cache attention over KIVI-packed low-bit K / V -- fp16 queries against a cache quantized by
 * lowbit_quant_pack_lastdim (new_pack.py:247-300; the prototype's driver attn_4bit_per_block.py:637-665):
 * replaces _quantized_flash_attn_forward / _fwd_kernel (src/triton/quantization/attn_4bit_per_block.py:28-553).
 *   q      fp16 [B,Nq,H,D], element strides (q_stride_b, q_stride_n, q_stride_h), last dim contiguous
 *   kcode  [B,D,H,N*bits/8] bytes (K packed along the sequence, per channel), kscale / kmn fp16 [B,D,H,N/group]
 *   vcode  [B,N,H,D*bits/8] bytes (V packed along the channels, per token),   vscale / vmn fp16 [B,N,H,D/group]
 *          all six contiguous; code i of a byte sits at bits [i*bits, (i+1)*bits)
 *   o      fp16, addressed like q; lse (may be NULL) f32 [B,H,lse_stride], natural log (m + log l, :372)
 * K^ = fma(code, scale, mn), V^ likewise, S = q.K^ in fp32, softmax(S * softmax_scale), o = P.V^ (:260-262, :330-372).
 * D in {64, 128}, bits in {4, 2}, group_size 32, N a multiple of 32; non-causal, no bias (the prototype's driver).
 * HBM-bound decode path: the cache is read once, key splits are merged through `workspace`
 * (>= lowbit_kv_attn_workspace_bytes()). */
int64_t lowbit_kv_attn_workspace_bytes(int B, int H, int Nq, int N, int D);
int lowbit_kv_attn_fwd(const void* q, const void* kcode, const void* kscale, const void* kmn, const void* vcode,
                       const void* vscale, const void* vmn, void* o, float* lse, void* workspace, int B, int H, int Nq,
                       int N, int D, int group_size, int bits, float softmax_scale, int64_t q_stride_b,
                       int64_t q_stride_n, int64_t q_stride_h, int64_t o_stride_b, int64_t o_stride_n,
                       int64_t o_stride_h, int lse_stride, void* stream);

/* Q6 -- V -> FP8 e4m3 per channel, transposed, padded, token-permuted
 * (src/quant.py:210-291; TransposePadPermuteKernel + MeanScaleKernel csrc/fused/fused.cu:263-428).
 * v8: e4m3 bytes addressed as [b][h][d][pos], pos contiguous in [0, Npad64), with byte strides
 * (v8_stride_b, v8_stride_h, v8_stride_d) -- the reference allocates [B,H,D,Npad] for HND and [B,D,H,Npad]
 * for NHD (src/quant.py:262-274); v_scale: f32 [B,H,D]; vm: NULL (smooth_v=False) or f32 [B,H,D].
 * Token r of every aligned 16-group lands at position (r/8)*2 + ((r/2)%4)*4 + (r%2) (fused.cu:290-292).
 * workspace: >= lowbit_v_fp8_workspace_bytes(). */
int64_t lowbit_v_fp8_workspace_bytes(int B, int H, int N, int D);
int lowbit_v_fp8_per_channel(const void* v, void* v8, float* v_scale, float* vm, void* workspace,
                             int B, int H, int N, int D,
                             int64_t stride_b, int64_t stride_h, int64_t stride_n,
                             int64_t v8_stride_b, int64_t v8_stride_h, int64_t v8_stride_d,
                             float scale_max, int dtype, void* stream);

/* sub_mean (src/quant.py:175-207; SubMeanKernel csrc/fused/fused.cu:200-261): out = fp16(v - vm), the difference
 * taken in the input dtype (__hsub2, :243) and then converted to fp16.  vm: [B,H,D] contiguous in v's dtype;
 * out: fp16, addressed by (out_stride_b, out_stride_h, out_stride_n) like v; head_dim any multiple of 8. */
int lowbit_sub_mean(const void* v, const void* vm, void* out, int B, int H, int N, int D,
                    int64_t stride_b, int64_t stride_h, int64_t stride_n,
                    int64_t out_stride_b, int64_t out_stride_h, int64_t out_stride_n, int dtype, void* stream);

/* E4 -- global max|x| of a tensor (compute_scale, src/core.py:1039-1047).  out: one f32 (device). */
int lowbit_abs_max(const void* x, float* out, int B, int H, int N, int D,
                   int64_t stride_b, int64_t stride_h, int64_t stride_n, int dtype, void* stream);

/* E4 -- global maximum and minimum of a tensor (the asymmetric branch of compute_scale, src/core.py:1043-1045:
 * `(max - min) / (2^bits - 1)`).  out: two f32 (device): out[0] = max, out[1] = min. */
int lowbit_min_max(const void* x, float* out, int B, int H, int N, int D,
                   int64_t stride_b, int64_t stride_h, int64_t stride_n, int dtype, void* stream);

/* A1/A2/A3 -- fused low-bit attention forward.
 * Replaces: _attn_fwd (src/triton/attn_qk_int8_per_block.py:24-167, host forward :169-238),
 *           _attn_fwd_base (src/triton/attn_qk_int8_per_block_causal.py:216-334),
 *           forward_merging (src/triton/quantization/attn_qk_int4_per_block.py:248-317) and the
 *           FP8-PV semantics of csrc/qattn/qk_int_sv_f8_cuda.cu:44-692.
 * q_codes int8 [B,Hq,Nq,D]; k_codes int8 [B,Hkv,Nk,D] (or packed, see qk_mode); v fp16 [B,Hkv,Nk,D]
 * (pv_mode F16) or e4m3 [b][h][d][pos] as produced by lowbit_v_fp8_per_channel (pv_mode E4M3: the three v strides
 * are then the byte strides of (b, h, d), positions contiguous, Npad64 of them);
 * q_scale f32 [B,Hq,ceil(Nq/128)], k_scale f32 [B,Hkv,ceil(Nk/64)] contiguous; v_scale/v_mean f32 [B,Hkv,D]
 * or NULL; kbits int32 [B,Hkv,ceil(Nk/64)] or NULL; o [B,Hq,Nq,D] of out_dtype; lse NULL or f32
 * [B,Hq,Nq] (base-2: log2(l)+m, as the reference kernel stores it).  D in {64,128}. */
int lowbit_attn_fwd(const void* q_codes, const void* k_codes, const void* v,
                    const float* q_scale, const float* k_scale,
                    const float* v_scale, const float* v_mean, const int32_t* kbits,
                    void* o, float* lse,
                    int B, int Hq, int Hkv, int Nq, int Nk, int D,
                    int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_n,
                    int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_n,
                    int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_n,
                    int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_n,
                    int qk_mode, int pv_mode, int out_dtype, int flags, void* stream);

/* Ring / sequence-parallel step: same contraction over one K/V shard, merged into un-normalised running state.
 * m_io, l_io: f32 [B,Hq,Nq]; o_acc_io: f32 [B,Hq,Nq,D] contiguous, all in true (dequantized) units:
 * O = o_acc / l relative to the running maximum m (base-2 log units).  q_offset / k_offset are the global token
 * positions of query row 0 / key 0 of the shards (causal masking across shards: key c is visible to row r iff
 * k_offset + c <= q_offset + r; a shard wholly in the future leaves the state untouched).  first != 0 initialises
 * the state instead of reading it.  Call lowbit_attn_finalize afterwards to produce o (and lse). */
int lowbit_attn_fwd_partial(const void* q_codes, const void* k_codes, const void* v,
                            const float* q_scale, const float* k_scale,
                            const float* v_scale, const float* v_mean, const int32_t* kbits,
                            float* m_io, float* l_io, float* o_acc_io,
                            int B, int Hq, int Hkv, int Nq, int Nk, int D,
                            int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_n,
                            int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_n,
                            int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_n,
                            int64_t q_offset, int64_t k_offset, int qk_mode, int pv_mode, int flags, int first,
                            void* stream);
int lowbit_attn_finalize(const float* m, const float* l, const float* o_acc, void* o, float* lse,
                         int B, int Hq, int Nq, int D,
                         int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_n,
                         int out_dtype, void* stream);

/* lse[b,h,n] = lse2/1.44269504 + (q . km) * sm_scale  -- the LSE fix-up of src/core.py:296-304,344-350. */
int lowbit_lse_fixup(float* lse, const void* q, const void* km, int B, int Hq, int Hkv, int Nq, int D,
                     int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_n,
                     float sm_scale, int dtype, void* stream);

/* The whole hot path in ONE call -- the host orchestration of src/core.py:194-352 (lowbit_fa_qk_int8_pv_fp16_triton,
 * :945-1036 for INT4 K): km = mean_n(k) (:293), Q codes with q_multiplier = sm_scale * log2(e) folded in and K codes of
 * k - km (:300-319, K blocks of 64, Q blocks of 128), attention forward (:321-341), and -- when lse != NULL -- the
 * natural-log lse with the K-smoothing correction (:343-350).  Launches exactly the kernels of lowbit_k_mean,
 * lowbit_quant_per_block (x2), lowbit_attn_fwd and lowbit_lse_fixup with the same arguments: bit-identical results, one
 * FFI call instead of five.  q, k: fp16 / bf16 (`dtype`), v: fp16, all [B,H,N,D] by strides; layout_nhd only fixes the
 * layout of the internal code tensors (that of their sources).  k_bits 8, or 4 with k_pack = 1 (two codes per byte).
 * workspace: lowbit_fa_fwd_workspace_bytes(...) bytes, 256-byte aligned; its contents are scratch.
 * flags: LOWBIT_ATTN_* as for lowbit_attn_fwd. */
int64_t lowbit_fa_fwd_workspace_bytes(int B, int Hq, int Hkv, int Nq, int Nk, int D, int k_bits, int k_pack);
int lowbit_fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, void* workspace,
                  int B, int Hq, int Hkv, int Nq, int Nk, int D, int layout_nhd,
                  int64_t q_stride_b, int64_t q_stride_h, int64_t q_stride_n,
                  int64_t k_stride_b, int64_t k_stride_h, int64_t k_stride_n,
                  int64_t v_stride_b, int64_t v_stride_h, int64_t v_stride_n,
                  int64_t o_stride_b, int64_t o_stride_h, int64_t o_stride_n,
                  float sm_scale, float q_multiplier, int k_bits, int k_pack, int smooth_k, int quant_mode, int dtype,
                  int out_dtype, int flags, void* stream);

/* diagnostics (not part of the reference surface): when set to a device buffer of 128*64 int32, the next
 * lowbit_attn_fwd launches make CTA (0,0,0) dump its raw int32 Q.K^T scores of key block 0. NULL disables. */
void lowbit_attn_set_debug_buffer(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* LOWBIT_FA_H_ */
