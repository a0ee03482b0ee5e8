"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the public Python API, which calls
the C ABI (liblowbit_fa_b200.so) -- never the oracle.  The oracle (oracle/*.py, CPU) and the golden vectors
(outputs of the reference's own Triton kernels, tests/golden/) are the checkers.

Bars: quantizer codes/scales and K mean BIT-EXACT; attention output within a stated tolerance of the reference
kernel's output (max-abs <= 4e-3 for fp16 outputs of O(1) magnitude, cos-sim >= 0.999 -- the north-star bar)
and cos-sim vs FP32 SDPA reported/asserted >= 0.999."""
import pytest
import torch

from conftest import cos_sim, golden_names, load_golden

pytestmark = pytest.mark.gpu

ATTN = golden_names("attn_")


@pytest.fixture(scope="module")
def L(cuda_dev):
    import lowbit_quant_fa2_paddle_b200 as pkg
    from lowbit_quant_fa2_paddle_b200 import _native
    _native.lib()  # fail loudly if the CUDA library is missing
    return pkg


def mk(b, h, n, d, layout, dtype, seed, bias=0.0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, h, n, d, generator=g) * scale
    if bias:
        x = x + bias * torch.randn(1, h, 1, d, generator=g)
    x = x.to(dtype)
    return x if layout == "HND" else x.permute(0, 2, 1, 3).contiguous()


# ------------------------------------------------------------------------------------------------ quantizers
@pytest.mark.parametrize("name", ATTN)
def test_quant_matches_reference_kernel_golden(L, cuda_dev, name):
    """Codes + scales bit-exact against the reference's own quant kernels (interpreter-run goldens)."""
    g = load_golden(name)
    q, k, km = g["q"].to(cuda_dev), g["k"].to(cuda_dev), g["km"].to(cuda_dev)
    assert torch.equal(L.k_mean(k, g["layout"]).cpu(), g["km"])
    qi, qs, ki, ks = L.per_block_int8(q, k, km=km, sm_scale=g["sm_scale"], tensor_layout=g["layout"])
    assert torch.equal(qi.cpu(), g["q_int8"]) and torch.equal(qs.cpu(), g["q_scale"])
    assert torch.equal(ki.cpu(), g["k_int8"]) and torch.equal(ks.cpu(), g["k_scale"])
    _, _, k4, k4s = L.per_block_q_int8_k_int4(q, k, km=km, sm_scale=g["sm_scale"], tensor_layout=g["layout"], pack=False)
    assert torch.equal(k4.cpu(), g["k_int4"]) and torch.equal(k4s.cpu(), g["k_int4_scale"])


@pytest.mark.parametrize("layout", ["HND", "NHD"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 2, 512, 64), (2, 3, 200, 128), (1, 1, 1, 64), (1, 2, 129, 128), (2, 2, 1111, 64)])
def test_quant_bit_exact_vs_oracle(L, cuda_dev, layout, dtype, shape):
    from oracle import quant as OQ
    b, h, n, d = shape
    q = mk(b, h, n, d, layout, dtype, 1)
    k = mk(b, h, n, d, layout, dtype, 2, bias=3.0)
    km_ref = OQ.k_mean(k, layout)
    km = L.k_mean(k.to(cuda_dev), layout)
    assert torch.equal(km.cpu(), km_ref)
    for backend, ofn in (("triton", OQ.per_block_int8_q1), ("cuda", OQ.per_block_int8_q2)):
        got = L.per_block_int8(q.to(cuda_dev), k.to(cuda_dev), km=km, tensor_layout=layout, backend=backend)
        ref = ofn(q, k, km_ref, tensor_layout=layout)
        for a, r, nm in zip(got, ref, ("q_int8", "q_scale", "k_int8", "k_scale")):
            assert torch.equal(a.cpu(), r), f"{backend} {nm}"
    # without smoothing
    got = L.per_block_int8(q.to(cuda_dev), k.to(cuda_dev), tensor_layout=layout, sm_scale=0.3)
    ref = OQ.per_block_int8_q1(q, k, None, sm_scale=0.3, tensor_layout=layout)
    assert all(torch.equal(a.cpu(), r) for a, r in zip(got, ref))
    # INT4 unpacked / packed, INT2 packed
    got = L.per_block_int4_unpack(q.to(cuda_dev), k.to(cuda_dev), km=km, tensor_layout=layout)
    ref = OQ.per_block_int8_q1(q, k, km_ref, tensor_layout=layout, kbits=4, qbits=4)
    assert all(torch.equal(a.cpu(), r) for a, r in zip(got, ref))
    got = L.per_block_q_int8_k_int4(q.to(cuda_dev), k.to(cuda_dev), km=km, tensor_layout=layout)
    ref = OQ.per_block_int8_q1(q, k, km_ref, tensor_layout=layout, kbits=4)
    assert torch.equal(got[2].cpu(), OQ.pack_codes(ref[2], 4)) and torch.equal(got[3].cpu(), ref[3])
    c2, s2 = L.per_block_k_lowbit(k.to(cuda_dev), km=km, bits=2, tensor_layout=layout)
    r2, rs2 = OQ.quant_per_block_q1(k - km_ref, 64, 1.0, layout, bits=2)
    assert torch.equal(c2.cpu(), OQ.pack_codes(r2, 2)) and torch.equal(s2.cpu(), rs2)


def test_quant_zero_block_and_extremes(L, cuda_dev):
    """All-zero block (reference: 0/0) -> scale 0, codes 0; fp16 extremes survive."""
    from oracle import quant as OQ
    q = mk(1, 1, 256, 64, "HND", torch.float16, 5)
    k = mk(1, 1, 256, 64, "HND", torch.float16, 6)
    k[:, :, 64:128] = 0
    k[0, 0, 200, 3] = 60000.0
    k[0, 0, 201, 4] = -6e-8
    got = L.per_block_int8(q.to(cuda_dev), k.to(cuda_dev))
    ref = OQ.per_block_int8_q1(q, k, None)
    assert all(torch.equal(a.cpu(), r) for a, r in zip(got, ref))
    assert float(got[3][0, 0, 1]) == 0.0 and int(got[2][0, 0, 64:128].abs().max()) == 0


def test_k_mean_exact_and_order_independent(L, cuda_dev):
    from oracle import quant as OQ
    k = (torch.randn(2, 3, 5000, 64) * 4 + 7).half()
    km = L.k_mean(k.to(cuda_dev))
    assert torch.equal(km.cpu(), OQ.k_mean(k))
    perm = torch.randperm(5000)
    assert torch.equal(L.k_mean(k[:, :, perm].contiguous().to(cuda_dev)).cpu(), km.cpu())


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("name", ATTN)
def test_attention_matches_reference_kernel_golden(L, cuda_dev, name):
    """Same codes/scales in, reference-kernel output out: max-abs <= 4e-3 (fp16) / 3.2e-2 (bf16), cos >= 0.999.
    Differences: P.V accumulates in fp32 here (fp16 per block in the reference), lazy rescale, ex2.approx."""
    g = load_golden(name)
    causal, layout = bool(g["causal"]), g["layout"]
    args = [g[k].to(cuda_dev) for k in ("q_int8", "k_int8")] + [g["v"].to(torch.float16).to(cuda_dev)] + \
           [g[k].to(cuda_dev) for k in ("q_scale", "k_scale")]
    fn = L.forward_causal if causal else L.forward
    kw = {} if causal else {"compat_tail": True}  # the reference does not mask tail keys
    o, lse2 = fn(*args, tensor_layout=layout, output_dtype=g["o"].dtype, return_lse=True, **kw)
    tol = 4e-3 if g["o"].dtype == torch.float16 else 3.2e-2
    err = (o.cpu().float() - g["o"].float()).abs().max().item()
    assert err <= tol, f"max-abs {err}"
    assert cos_sim(o.cpu(), g["o"]) >= 0.999
    assert (lse2.cpu() - g["lse2"]).abs().max() <= 2e-3


@pytest.mark.parametrize("case", [
    # b, hq, hkv, n, d, layout, causal, dtype
    (1, 2, 2, 512, 64, "HND", False, torch.float16),   # BASELINE config 1
    (1, 2, 2, 512, 64, "HND", True, torch.float16),
    (1, 2, 2, 512, 128, "HND", True, torch.float16),
    (2, 4, 2, 384, 128, "NHD", False, torch.float16),  # GQA + NHD
    (1, 4, 1, 320, 64, "NHD", True, torch.bfloat16),
    (1, 2, 2, 200, 64, "HND", False, torch.float16),   # ragged tails
    (1, 2, 2, 77, 128, "HND", True, torch.float16),
    (1, 1, 1, 1, 64, "HND", False, torch.float16),
    (1, 2, 2, 129, 64, "NHD", False, torch.float16),
    (1, 2, 2, 96, 32, "HND", False, torch.float16),    # head_dim padded 32 -> 64
    (1, 2, 2, 160, 96, "NHD", True, torch.float16),    # head_dim padded 96 -> 128
])
@pytest.mark.parametrize("smooth_k", [True, False])
def test_api_vs_oracle_and_sdpa(L, cuda_dev, case, smooth_k):
    from oracle import attention as OA
    b, hq, hkv, n, d, layout, causal, dtype = case
    q = mk(b, hq, n, d, layout, dtype, 11)
    k = mk(b, hkv, n, d, layout, dtype, 12, bias=4.0)  # channel bias: smoothing matters
    v = mk(b, hkv, n, d, layout, dtype, 13)
    o, lse = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                                is_causal=causal, smooth_k=smooth_k, return_lse=True)
    assert o.shape == q.shape and o.dtype == dtype and lse.shape == (b, hq, n)
    oref, lref = OA.lowbit_fa_api(q, k, v, layout, causal, smooth_k=smooth_k, return_lse=True, compat_tail=False,
                                  pv_accum="fp32")
    tol = 4e-3 if dtype == torch.float16 else 3.2e-2
    assert (o.cpu().float() - oref.float()).abs().max() <= tol
    assert (lse.cpu() - lref).abs().max() <= 1e-2
    sd, lsd = OA.sdpa_fp32(q, k, v, layout, causal, return_lse=True)
    if smooth_k:
        assert cos_sim(o.cpu(), sd) >= 0.999
        assert (lse.cpu() - lsd).abs().max() <= 0.1
    o_only = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                                is_causal=causal, smooth_k=smooth_k)
    assert torch.equal(o_only, o)


def test_cuda_rounding_backend(L, cuda_dev):
    from oracle import attention as OA
    q, k, v = (mk(1, 2, 256, 64, "HND", torch.float16, s) for s in (1, 2, 3))
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), quantization_backend="cuda")
    assert cos_sim(o.cpu(), OA.sdpa_fp32(q, k, v)) >= 0.999


def test_tail_masking_modes(L, cuda_dev):
    """Nk % 64 != 0: default masks the tail keys (matches SDPA); compat_tail reproduces the reference's
    phantom zero-score keys (matches the emulator with compat_tail=True)."""
    from oracle import attention as OA
    from oracle import quant as OQ
    q, k, v = (mk(1, 2, 200, 64, "HND", torch.float16, s) for s in (21, 22, 23))
    qi, qs, ki, ks = OQ.per_block_int8_q1(q, k, None)
    dev = [t.to(cuda_dev) for t in (qi, ki, v, qs, ks)]
    for compat in (False, True):
        o, _ = L.forward(*dev, compat_tail=compat)
        ref, _ = OA.attn_block_emulator(qi, ki, v, qs, ks, compat_tail=compat, pv_accum="fp32")
        other, _ = OA.attn_block_emulator(qi, ki, v, qs, ks, compat_tail=not compat, pv_accum="fp32")
        assert (o.cpu().float() - ref.float()).abs().max() <= 2e-3
        assert (o.cpu().float() - other.float()).abs().max() > (o.cpu().float() - ref.float()).abs().max()


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_c2(L, cuda_dev):
    """BASELINE config 2 (B4 H32 N4096 D64): size-independent properties.
    (1) rows of softmax sum to one: V = per-channel constant -> O = that constant (up to fp16 rounding);
    (2) batch/head independence + determinism: computing one (b,h) slice alone is bit-identical;
    (3) HND == NHD bit-identical; (4) codes bit-exact vs oracle on a slice; (5) cos-sim vs FP32 SDPA on a slice."""
    from oracle import attention as OA
    from oracle import quant as OQ
    torch.manual_seed(0)
    b, h, n, d = 4, 32, 4096, 64
    q = torch.randn(b, h, n, d, dtype=torch.float16, device=cuda_dev)
    k = torch.randn(b, h, n, d, dtype=torch.float16, device=cuda_dev) + 2.0
    v = torch.randn(b, h, n, d, dtype=torch.float16, device=cuda_dev)
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v)
    assert not torch.isnan(o).any()
    # (1)
    c = torch.linspace(-2, 2, d, device=cuda_dev).half()
    oc = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, c.expand(b, h, n, d).contiguous())
    assert (oc.float() - c.float()).abs().max() <= 4e-3
    # (2)
    o_slice = L.lowbit_fa_qk_int8_pv_fp16_triton(q[1:2, 5:6].contiguous(), k[1:2, 5:6].contiguous(), v[1:2, 5:6].contiguous())
    assert torch.equal(o_slice[0, 0], o[1, 5])
    assert torch.equal(L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v), o)
    # (3)
    qn, kn, vn = (t.permute(0, 2, 1, 3).contiguous() for t in (q, k, v))
    on = L.lowbit_fa_qk_int8_pv_fp16_triton(qn, kn, vn, tensor_layout="NHD")
    assert torch.equal(on.permute(0, 2, 1, 3), o)
    # (4) + (5) on one (b, h) slice
    qs_, ks_, vs_ = (t[2:3, 7:8].cpu() for t in (q, k, v))
    km = L.k_mean(k)
    got = L.per_block_int8(q, k, km=km)
    ref = OQ.per_block_int8_q1(qs_, ks_, OQ.k_mean(ks_))
    assert torch.equal(got[0][2:3, 7:8].cpu(), ref[0]) and torch.equal(got[2][2:3, 7:8].cpu(), ref[2])
    assert torch.equal(got[1][2:3, 7:8].cpu(), ref[1]) and torch.equal(got[3][2:3, 7:8].cpu(), ref[3])
    sd = OA.sdpa_fp32(qs_, ks_, vs_)
    assert cos_sim(o[2:3, 7:8].cpu(), sd) >= 0.999
    # causal at full size: first row attends only to key 0 -> O[0] == V[0]
    ocz = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v, is_causal=True)
    assert (ocz[:, :, 0].float() - v[:, :, 0].float()).abs().max() <= 2e-3
    sdc = OA.sdpa_fp32(qs_, ks_, vs_, is_causal=True)
    assert cos_sim(ocz[2:3, 7:8].cpu(), sdc) >= 0.999


def test_large_magnitude_inputs(L, cuda_dev):
    """The reference kernel bench uses randint(-100,100) Q/K without smoothing (utils/benchmark.py:215-230):
    near one-hot softmax, large score range."""
    from oracle import attention as OA
    g = torch.Generator().manual_seed(3)
    q = torch.randint(-100, 100, (1, 2, 512, 64), generator=g).half()
    k = torch.randint(-100, 100, (1, 2, 512, 64), generator=g).half()
    v = torch.randn(1, 2, 512, 64, generator=g).half()
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), smooth_k=False)
    oref = OA.lowbit_fa_api(q, k, v, smooth_k=False, compat_tail=False, pv_accum="fp32")
    assert not torch.isnan(o).any()
    assert (o.cpu().float() - oref.float()).abs().max() <= 1e-2
    assert cos_sim(o.cpu(), oref) >= 0.999


def test_multi_precision_and_select(L, cuda_dev):
    q, k, v = (mk(1, 2, 256, 64, "HND", torch.float16, s).to(cuda_dev) for s in (31, 32, 33))
    kind = L.select_quantization(q, k, v)
    sc = float(L.compute_scale(q))
    assert abs(sc - float(q.float().abs().max()) / 127) < 1e-6
    avg = sum(float(t.float().abs().max()) / 127 for t in (q, k, v)) / 3
    assert kind == ("FP16" if avg > 0.2 else "INT8" if avg > 0.05 else "INT4")
    o = L.lowbit_fa_multi_precision(q, k, v)
    assert o.shape == q.shape and not torch.isnan(o).any()


@pytest.mark.parametrize("shape,dtype,layout", [((2, 3, 300, 64), torch.float16, "HND"), ((1, 2, 77, 128), torch.bfloat16, "HND"),
                                                ((2, 129, 2, 128), torch.float16, "NHD"), ((1, 2, 50, 40), torch.float16, "HND"),
                                                ((1, 1, 9, 96), torch.bfloat16, "HND")])
@pytest.mark.parametrize("shift", [0.0, 3.0, -5.0])
def test_compute_scale_asymmetric(L, cuda_dev, shape, dtype, layout, shift):
    """compute_scale(symmetric=False) = (max - min) / (2^bits - 1) (core.py:1043-1045): exact extremes from the
    lowbit_min_max kernel (all-positive, all-negative and mixed tensors; head dims that need padding)."""
    g = torch.Generator().manual_seed(hash((shape, shift)) % 1000)
    x = (torch.randn(shape, generator=g) + shift).to(dtype).to(cuda_dev)
    from lowbit_quant_fa2_paddle_b200 import quant as Qz
    if shape[-1] in (64, 128):
        mm = Qz.min_max(x, layout).cpu()
        assert mm[0].item() == x.float().max().item() and mm[1].item() == x.float().min().item()
    for bits in (8, 4):
        sc = float(L.compute_scale(x, bits=bits, symmetric=False, tensor_layout=layout))
        ref = (x.float().max().item() - x.float().min().item()) / (2 ** bits - 1)
        assert abs(sc - ref) <= 1e-6 * max(1.0, abs(ref))


# ------------------------------------------------------------------------------------------------ Q3 / Q5 / Q6
@pytest.mark.parametrize("name", [n for n in ATTN if "causal" not in n])
@pytest.mark.parametrize("bits", [8, 4])
def test_per_thread_quant_matches_reference_golden(L, cuda_dev, name, bits):
    """Q3: codes + scales bit-exact against the reference's per-thread Triton kernels."""
    g = load_golden(name)
    fn = L.per_thread_int8 if bits == 8 else L.per_thread_int4
    qi, qs, ki, ks = fn(g["q"].to(cuda_dev), g["k"].to(cuda_dev), km=g["km"].to(cuda_dev), tensor_layout=g["layout"])
    p = f"pt{bits}_"
    assert torch.equal(qi.cpu(), g[p + "q"]) and torch.equal(qs.cpu(), g[p + "qs"])
    assert torch.equal(ki.cpu(), g[p + "k"]) and torch.equal(ks.cpu(), g[p + "ks"])


def test_per_warp_quant_vs_oracle(L, cuda_dev):
    from oracle import quant as OQ
    for layout, (b, h, n, d) in (("HND", (1, 2, 300, 64)), ("NHD", (2, 2, 130, 128))):
        q = mk(b, h, n, d, layout, torch.float16, 41)
        k = mk(b, h, n, d, layout, torch.float16, 42, bias=2.0)
        km = OQ.k_mean(k, layout)
        got = L.per_warp_int8(q.to(cuda_dev), k.to(cuda_dev), km=km.to(cuda_dev), tensor_layout=layout)
        ref = OQ.per_warp_int8_q2(q, k, km, tensor_layout=layout)
        for a, r in zip(got, ref):
            assert torch.equal(a.cpu(), r)


@pytest.mark.parametrize("bit", [2, 4, 8])
def test_kivi_pack_matches_reference_golden(L, cuda_dev, bit):
    """Q5: packed codes, scales and zero points bit-exact against the reference's min/max + pack kernels."""
    from oracle import quant as OQ
    g = load_golden(f"kivi_b{bit}")
    code, scale, mn = L.triton_quantize_and_pack_along_last_dim(g["data"].to(cuda_dev), 32, bit)
    assert torch.equal(code.cpu(), g["code"])
    assert torch.equal(scale.cpu().view(torch.int16), g["scale"].view(torch.int16))
    assert torch.equal(mn.cpu().view(torch.int16), g["mn"].view(torch.int16))
    big = (torch.randn(2, 64, 4, 256) * 3).half()
    big[0, 0, 0, :32] = 1.5  # constant group
    c2, s2, m2 = L.triton_quantize_and_pack_along_last_dim(big.to(cuda_dev), 32, bit)
    rc, rs, rm = OQ.kivi_quantize_and_pack(big, 32, bit)
    assert torch.equal(c2.cpu(), rc) and torch.equal(s2.cpu().view(torch.int16), rs.view(torch.int16))
    assert torch.equal(m2.cpu().view(torch.int16), rm.view(torch.int16))


@pytest.mark.parametrize("layout", ["HND", "NHD"])
@pytest.mark.parametrize("smooth_v", [True, False])
@pytest.mark.parametrize("shape,dtype", [((1, 2, 200, 64), torch.float16), ((2, 3, 77, 128), torch.bfloat16),
                                         ((1, 2, 1024, 128), torch.float16)])
def test_v_fp8_per_channel_vs_oracle(L, cuda_dev, layout, smooth_v, shape, dtype):
    """Q6: e4m3 bytes, scales and means bit-exact against the IEEE restatement of fused.cu (the restatement itself is
    pinned to the reference's own fused.cu kernels by tests/test_fused_golden.py)."""
    from oracle import quant as OQ
    b, h, n, d = shape
    v = mk(b, h, n, d, layout, dtype, 51, bias=1.0)
    v8, vs, vm = L.per_channel_fp8(v.to(cuda_dev), tensor_layout=layout, smooth_v=smooth_v)
    r8, rs, rm = OQ.per_channel_fp8(v, tensor_layout=layout, smooth_v=smooth_v)
    assert v8.shape == r8.shape and v8.dtype == torch.float8_e4m3fn
    assert torch.equal(vs.cpu(), rs)
    if smooth_v:
        assert torch.equal(vm.cpu(), rm)
    else:
        assert vm is None
    assert torch.equal(v8.cpu().view(torch.uint8), r8.view(torch.uint8))


# ------------------------------------------------------------------------------------------------ INT4-K / FP8-PV / ring
K4_CASES = [
    # b, hq, hkv, n, d, layout, causal
    (1, 2, 2, 512, 64, "HND", False),
    (1, 2, 2, 512, 64, "HND", True),
    (2, 4, 2, 384, 128, "NHD", False),
    (1, 2, 1, 320, 128, "HND", True),
    (1, 2, 2, 200, 64, "NHD", False),   # ragged tail
    (1, 2, 2, 77, 128, "HND", True),
    (1, 1, 1, 1, 64, "HND", False),
    (1, 1, 1, 33, 64, "HND", True),
    (1, 3, 3, 1000, 64, "NHD", True),
]


@pytest.mark.parametrize("case", K4_CASES)
def test_packed_int4_k_matches_unpacked(L, cuda_dev, case):
    """qk_mode Q8K4 (K packed two codes per byte in HBM, expanded to code*16 in shared memory, Q tile permuted,
    1/16 folded into the scale) against the same codes fed one per int8: the integer scores are identical.  The
    scores leave TMEM as fp32 1.5 * 2^23 + s; for fine scales the constant goes out through the addend of the scaling
    FMA, nm - 12582912 * scale, whose rounding depends on the scale (scale / 16 for the packed operand), so the two
    forms' exponents differ by <= 2^-11, i.e. P by less than its own fp16 rounding (the unpacked INT4 codes have a
    coarse scale and take the exact-subtraction path): O within 2 fp16 ulps of its magnitude, lse within 1e-3."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    b, hq, hkv, n, d, layout, causal = case
    q = mk(b, hq, n, d, layout, torch.float16, 41).to(cuda_dev)
    k = mk(b, hkv, n, d, layout, torch.float16, 42, bias=2.0).to(cuda_dev)
    v = mk(b, hkv, n, d, layout, torch.float16, 43).to(cuda_dev)
    km = L.k_mean(k, layout)
    qc, qs, k4, ks = L.per_block_q_int8_k_int4(q, k, km=km, tensor_layout=layout, pack=False)
    _, _, k4p, ksp = L.per_block_q_int8_k_int4(q, k, km=km, tensor_layout=layout, pack=True)
    assert torch.equal(ks, ksp) and k4p.shape[-1] == d // 2
    fn = L.forward_causal if causal else L.forward
    o_u, lse_u = fn(qc, k4, v, qs, ks, tensor_layout=layout, return_lse=True)
    o_p, lse_p = fn(qc, k4p, v, qs, ks, tensor_layout=layout, return_lse=True, qk_mode=NV.QK_Q8K4)
    tol = 2.0 * 2.0 ** -10 * max(o_u.float().abs().max().item(), 2.0 ** -10)
    assert (o_u.float() - o_p.float()).abs().max().item() <= tol
    assert (lse_u - lse_p).abs().max().item() <= 1e-3


@pytest.mark.parametrize("entry", ["int4", "q8k4"])
@pytest.mark.parametrize("case", K4_CASES[:6])
def test_int4_api_vs_oracle_and_sdpa(L, cuda_dev, case, entry):
    """E2 / E3 entry points (packed INT4 K path) against the oracle's Q INT8 / K INT4 per-block semantics
    (reference kernel parity unpinned) and FP32 SDPA."""
    from oracle import attention as OA
    b, hq, hkv, n, d, layout, causal = case
    q = mk(b, hq, n, d, layout, torch.float16, 51)
    k = mk(b, hkv, n, d, layout, torch.float16, 52, bias=3.0)
    v = mk(b, hkv, n, d, layout, torch.float16, 53)
    fn = L.lowbit_fa_qk_int4_pv_fp16_triton if entry == "int4" else L.lowbit_fa_q_int8_k_int4_pv_fp16
    o, lse = fn(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout, is_causal=causal, return_lse=True)
    oref, lref = OA.lowbit_fa_api(q, k, v, layout, causal, return_lse=True, compat_tail=False, pv_accum="fp32", qk="int4")
    assert (o.cpu().float() - oref.float()).abs().max() <= 4e-3
    assert (lse.cpu() - lref).abs().max() <= 1e-2
    if n >= 64:
        assert cos_sim(o.cpu(), OA.sdpa_fp32(q, k, v, layout, causal)) >= 0.98  # 4-bit K codes: format error, not kernel error


FP8_CASES = [
    # b, hq, hkv, n, d, layout, causal, dtype, qk
    (1, 2, 2, 512, 64, "HND", False, torch.float16, "int8"),
    (1, 2, 2, 512, 128, "HND", True, torch.float16, "int4"),   # BASELINE config 3 combination
    (2, 4, 2, 384, 128, "NHD", False, torch.float16, "int8"),
    (1, 4, 1, 320, 64, "NHD", True, torch.bfloat16, "int4"),
    (1, 2, 2, 200, 64, "HND", False, torch.float16, "int4"),
    (1, 2, 2, 77, 128, "NHD", True, torch.float16, "int8"),
    (1, 1, 1, 1, 64, "HND", False, torch.float16, "int8"),
    (1, 2, 2, 1030, 128, "HND", False, torch.float16, "int4"),
]


@pytest.mark.parametrize("smooth_v", [False, True])
@pytest.mark.parametrize("case", FP8_CASES)
def test_fp8_pv_vs_oracle_and_sdpa(L, cuda_dev, case, smooth_v):
    """A3 semantics (P~ = e4m3(exp2(s-m+off)), denominator over the rounded P~, epilogue * v_scale + v_mean) on
    tcgen05 kind::f8f6f4.  The kernel keeps a lazy reference maximum (P~ max in [112, 448]) where the spec uses the
    exact running maximum (P~ max = 448): same algorithm, different e4m3 rounding grid, so the comparison with
    the oracle is by tolerance: max-abs <= 0.05 (e4m3 has 3 mantissa bits), cos >= 0.999; vs FP32 SDPA cos >= 0.998."""
    from oracle import attention as OA
    b, hq, hkv, n, d, layout, causal, dtype, qk = case
    q = mk(b, hq, n, d, layout, dtype, 61)
    k = mk(b, hkv, n, d, layout, dtype, 62, bias=3.0)
    v = mk(b, hkv, n, d, layout, dtype, 63, bias=1.0 if smooth_v else 0.0)
    if qk == "int8":
        o, lse = L.lowbit_fa_qk_int8_pv_fp8_cuda(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                                 is_causal=causal, pv_accum_dtype="fp32", smooth_v=smooth_v,
                                                 return_lse=True)
    else:
        o, lse = L.lowbit_fa_qk_int4_pv_fp8(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                            is_causal=causal, smooth_v=smooth_v, return_lse=True)
    assert o.shape == q.shape and o.dtype == dtype and not torch.isnan(o).any()
    oref, lref = OA.lowbit_fa_api(q, k, v, layout, causal, return_lse=True, compat_tail=False, qk=qk, pv="fp8",
                                  smooth_v=smooth_v)
    err = (o.cpu().float() - oref.float()).abs().max().item()
    assert err <= 0.05 * max(1.0, float(v.float().abs().max()) / 4), f"max-abs {err}"
    assert (lse.cpu() - lref).abs().max() <= 3e-2
    if n >= 64:
        assert cos_sim(o.cpu(), oref) >= 0.999
        assert cos_sim(o.cpu(), OA.sdpa_fp32(q, k, v, layout, causal)) >= (0.998 if qk == "int8" else 0.98)


def test_fp8_pv_constant_v_is_exact(L, cuda_dev):
    """Rows of P~ / sum(P~) sum to one by construction (the denominator is taken over the rounded P~): a V that is
    constant per channel must come back as that constant up to e4m3 rounding of V itself and f16 partial sums."""
    b, h, n, d = 1, 2, 1024, 64
    q = mk(b, h, n, d, "HND", torch.float16, 71).to(cuda_dev)
    k = mk(b, h, n, d, "HND", torch.float16, 72).to(cuda_dev)
    c = torch.linspace(-2, 2, d).half()
    v = c.expand(b, h, n, d).contiguous().to(cuda_dev)
    o = L.lowbit_fa_qk_int8_pv_fp8_cuda(q, k, v, pv_accum_dtype="fp32")
    assert (o.cpu().float() - c.float()).abs().max() <= 5e-3  # per-channel scale makes a constant channel exact in e4m3


@pytest.mark.parametrize("cfg", [
    # d, layout, causal, qk, pv, nshard
    (64, "HND", False, "int8", "fp16", 4),
    (64, "NHD", True, "int8", "fp16", 4),
    (128, "HND", True, "int4", "fp16", 2),
    (128, "NHD", False, "int4", "fp8", 4),
    (64, "HND", True, "int4", "fp8", 8),
    (128, "HND", True, "int8", "fp8-shard", 4),
])
def test_ring_partial_steps_match_single_pass(L, cuda_dev, cfg):
    """Sequence-parallel decomposition: Q and K/V split into P shards along N; every Q shard visits every K/V
    shard in ring order through lowbit_attn_fwd_partial (causal masking by global offsets, shards wholly in the
    future are no-ops, including as the first step) and lowbit_attn_finalize.  With 64-aligned shards and a global
    km the per-shard codes/scales equal the single-pass ones, so the result must match the single-pass kernel
    up to fp32 merge rounding (<= 2e-3).  FP8 P.V: "fp8" slices one globally quantized V^T (same e4m3 codes as
    the single pass; P~ rounding grids still differ because the lazy maximum evolves differently: <= 0.05);
    "fp8-shard" quantizes every V shard with its own per-channel scale, as ranks of a real ring do, so V codes
    differ by up to one e4m3 step (2^-3 of the value) from the single pass: |diff| <= 0.125 max|v|."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    from lowbit_quant_fa2_paddle_b200 import attention as A
    from lowbit_quant_fa2_paddle_b200 import quant as Qz
    d, layout, causal, qk, pv, P = cfg
    b, hq, hkv, n = 1, 4, 2, 1024
    ns = n // P
    q = mk(b, hq, n, d, layout, torch.float16, 81).to(cuda_dev)
    k = mk(b, hkv, n, d, layout, torch.float16, 82, bias=2.0).to(cuda_dev)
    v = mk(b, hkv, n, d, layout, torch.float16, 83).to(cuda_dev)
    km = L.k_mean(k, layout)
    kbits, packed = (8, False) if qk == "int8" else (4, True)
    qk_mode = NV.QK_Q8K4 if packed else NV.QK_I8
    vq_shard = pv == "fp8-shard"
    pv = "fp8" if vq_shard else pv
    pv_mode = NV.PV_E4M3 if pv == "fp8" else NV.PV_F16
    sm = d ** -0.5
    seq = 2 if layout == "HND" else 1
    sh = lambda t, i: t.narrow(seq, i * ns, ns)  # strided views: the kernels take strides

    def quant(qx, kx):
        return Qz._per_block(qx, kx, km, 128, 64, sm, layout, 8, kbits, packed, "triton")

    def vprep(vx):
        if pv == "fp8":
            return Qz.per_channel_fp8(vx, layout, smooth_v=False)
        return vx, None, None

    # single pass
    qc, qs, kc, ks = quant(q, k)
    v1, vs1, vm1 = vprep(v)
    o_ref, lse_ref = A._forward(qc, kc, v1, qs, ks, layout, torch.float16, True, causal, qk_mode=qk_mode,
                                pv_mode=pv_mode, v_scale=vs1, v_mean=vm1)
    # ring
    shards = []
    for i in range(P):
        qci, qsi, kci, ksi = quant(sh(q, i), sh(k, i))
        if pv == "fp8" and not vq_shard:  # positions [i*ns, (i+1)*ns) of the global V^T (64-aligned: permutation-safe)
            vi, vsi, vmi = v1.narrow(3, i * ns, ns), vs1, vm1
        else:
            vi, vsi, vmi = vprep(sh(v, i))
        shards.append((qci, qsi, kci, ksi, vi, vsi, vmi))
    outs, lses = [], []
    for r in range(P):
        st = None
        for step in range(P):
            s = (r + 1 + step) % P  # start with the NEXT rank's shard: for causal that is a future (no-op) shard
            qci, qsi = shards[r][0], shards[r][1]
            _, _, kci, ksi, vi, vsi, vmi = shards[s]
            st = A.forward_partial(st, qci, kci, vi, qsi, ksi, layout, causal=causal, q_offset=r * ns, k_offset=s * ns,
                                   qk_mode=qk_mode, pv_mode=pv_mode, v_scale=vsi, v_mean=vmi)
        o_r, lse_r = A.finalize(st, shards[r][0], layout, torch.float16, return_lse=True)
        outs.append(o_r)
        lses.append(lse_r)
    o_ring = torch.cat(outs, dim=seq)
    lse_ring = torch.cat(lses, dim=2)
    diff = (o_ring.float() - o_ref.float()).abs()
    if vq_shard:
        assert diff.max() <= 0.125 * float(v.float().abs().max())
    else:
        assert diff.max() <= (2e-3 if pv == "fp16" else 0.05)
    assert (lse_ring - lse_ref).abs().max() <= (2e-3 if pv == "fp16" else 3e-2)
    assert cos_sim(o_ring.cpu(), o_ref.cpu()) >= (0.9999 if pv == "fp16" else 0.999)


def test_ring_attention_world1_cuda_backend(L, cuda_dev):
    """parallel.ring_attention with the product CudaBackend on one GPU (world_size 1, two zig-zag chunks): exercises
    quantization into the flat message views, the per-chunk partial/finalize calls and the LSE fix-up."""
    import os
    import socket
    import torch.distributed as dist
    from lowbit_quant_fa2_paddle_b200 import parallel as P
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda_dev)
    try:
        for layout, causal, qk, pv, d in (("HND", True, "int4", "fp16", 64), ("NHD", False, "int8", "fp16", 128),
                                          ("HND", True, "int4", "fp8", 128), ("HND", True, "mixed", "fp16", 128),
                                          ("NHD", False, "mixed", "fp16", 64)):
            q = mk(1, 4, 1024, d, layout, torch.float16, 91).to(cuda_dev)
            k = mk(1, 2, 1024, d, layout, torch.float16, 92, bias=2.0).to(cuda_dev)
            if qk == "mixed":
                k = _mixed_k(1, 2, 1024, d, layout, torch.float16, 92).to(cuda_dev)
            v = mk(1, 2, 1024, d, layout, torch.float16, 93).to(cuda_dev)
            o, lse = P.ring_attention(q, k, v, tensor_layout=layout, is_causal=causal, qk=qk, pv=pv, return_lse=True)
            if qk == "mixed":
                fn = L.lowbit_fa_q_int8_k_dynamic
            elif pv == "fp8":
                fn = L.lowbit_fa_qk_int4_pv_fp8
            else:
                fn = L.lowbit_fa_qk_int4_pv_fp16_triton if qk == "int4" else L.lowbit_fa_qk_int8_pv_fp16_triton
            o1, lse1 = fn(q, k, v, tensor_layout=layout, is_causal=causal, return_lse=True)
            tol = 4e-3 if pv == "fp16" else 0.125 * float(v.abs().max())
            assert (o.float() - o1.float()).abs().max() <= tol
            assert (lse - lse1).abs().max() <= (1e-2 if pv == "fp16" else 5e-2)
    finally:
        dist.destroy_process_group()


def test_fast_exact_division_equals_ieee_division(L, cuda_dev):
    """The Q1 quantizer takes RN(x/scale) from reciprocal + exact FMA remainder + corrected FMA (3 instructions)
    instead of the IEEE division (~25).  Both must give identical codes and scales: checked on every finite fp16
    value against many block maxima and sm_scale factors (~2e8 (x, scale) pairs), on bf16 data spanning the fp32
    exponent range, and on INT4 / INT2 targets."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    from lowbit_quant_fa2_paddle_b200 import quant as Qz
    allh = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    allh = allh[torch.isfinite(allh)]                        # 63488 finite fp16 values
    g = torch.Generator().manual_seed(5)
    reps = 32
    x = allh.repeat(reps)[torch.randperm(allh.numel() * reps, generator=g)]
    n = x.numel() // 64 // 128 * 128
    x = x[: n * 64].view(1, 1, n, 64).clone()
    # give the 128-row blocks different magnitudes so the per-block scales differ
    mag = torch.exp2(torch.randint(-14, 1, (n // 128,), generator=g).float()).repeat_interleave(128).view(1, 1, n, 1)
    x = (x.float() * mag).clamp(-65504, 65504).half().to(cuda_dev)
    for sm in (1.0, 0.18033688, 0.7071, 3.3333, 1e-3):
        for bits in (8, 4, 2):
            for blk in (128, 64):
                fast = Qz._quant_one(x, None, blk, bits, False, sm, NV.QMODE_TRITON, "HND")
                slow = Qz._quant_one(x, None, blk, bits, False, sm, NV.QMODE_TRITON | NV.QMODE_FLAG_IEEE_DIV, "HND")
                assert torch.equal(fast[0], slow[0]) and torch.equal(fast[1], slow[1]), (sm, bits, blk)
    xb = (torch.randn(1, 2, 4096, 128, generator=g) * torch.exp2(torch.randint(-120, 120, (1, 2, 64, 1), generator=g).float()
          ).repeat_interleave(64, dim=2)).bfloat16().to(cuda_dev)
    fast = Qz._quant_one(xb, None, 64, 8, False, 1.0, NV.QMODE_TRITON, "HND")
    slow = Qz._quant_one(xb, None, 64, 8, False, 1.0, NV.QMODE_TRITON | NV.QMODE_FLAG_IEEE_DIV, "HND")
    assert torch.equal(fast[0], slow[0]) and torch.equal(fast[1], slow[1])


def test_k_mean_exact_on_all_fp16_values(L, cuda_dev):
    """The split hi/lo int32 accumulation of k_mean is exact: a column holding every finite fp16 value (sum 0) next
    to columns with known sums, N large enough to cross the 256-row flush."""
    from oracle import quant as OQ
    allh = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    allh = allh[torch.isfinite(allh)]
    g = torch.Generator().manual_seed(6)
    n = allh.numel()
    cols = [allh[torch.randperm(n, generator=g)] for _ in range(60)]
    cols += [torch.full((n,), 65504.0).half(), torch.full((n,), -65504.0).half(),
             torch.full((n,), 6e-8).half(), allh.abs()]
    k = torch.stack(cols, dim=1).view(1, 1, n, 64).contiguous()
    km = L.k_mean(k.to(cuda_dev)).cpu()
    assert torch.equal(km, OQ.k_mean(k))
    assert torch.equal(km.view(-1)[:60], torch.zeros(60).half())


# ------------------------------------------------------------------------------------------------ fused K smoothing
@pytest.mark.parametrize("layout", ["HND", "NHD"])
@pytest.mark.parametrize("b,h,n,d", [(1, 2, 1, 64), (2, 3, 63, 64), (1, 2, 64, 128), (2, 2, 65, 64), (1, 3, 512, 64),
                                     (2, 2, 700, 128), (1, 4, 4096, 64), (1, 2, 5000, 128), (1, 1, 12000, 64)])
def test_k_smooth_quant_cluster_kernel_is_bit_identical(L, cuda_dev, layout, b, h, n, d):
    """lowbit_k_smooth_quant (one cluster launch, K read once, slice parked in shared memory, column sums over DSMEM)
    against lowbit_k_mean + lowbit_quant_per_block: km, codes and scales bit for bit, for int8 / packed INT4 / INT2 and
    both rounding conventions; and against the CPU oracle for the default mode."""
    from oracle import quant as OQ
    k = mk(b, h, n, d, layout, torch.float16, 71, bias=2.5)
    dk = k.to(cuda_dev)
    assert L.k_smooth_quant_supported(dk, layout)
    km_ref = L.k_mean(dk, layout)
    for bits, pack, backend in ((8, False, "triton"), (4, True, "triton"), (2, True, "triton"), (8, False, "cuda"),
                                (4, False, "triton_gpu")):
        km, kc, ks = L.k_smooth_quant(dk, bits=bits, pack=pack, tensor_layout=layout, backend=backend)
        from lowbit_quant_fa2_paddle_b200 import quant as Qz
        kc_ref, ks_ref = Qz._quant_one(dk, km_ref, 64, bits, pack, 1.0, Qz._MODES[backend], layout)
        assert torch.equal(km, km_ref), (bits, pack, backend)
        assert torch.equal(kc, kc_ref), (bits, pack, backend)
        assert torch.equal(ks.view(torch.int32), ks_ref.view(torch.int32)), (bits, pack, backend)
    km, kc, ks = L.k_smooth_quant(dk, tensor_layout=layout)
    km_o = OQ.k_mean(k, layout)
    _, _, kc_o, ks_o = OQ.per_block_int8_q1(k, k, km_o, tensor_layout=layout)
    assert torch.equal(km.cpu(), km_o) and torch.equal(kc.cpu(), kc_o) and torch.equal(ks.cpu(), ks_o)


def test_k_smooth_quant_unsupported_shapes_fall_back(L, cuda_dev):
    """bf16 and slices larger than a cluster's shared memory are not taken by the fused kernel: the C entry point says
    so, and the operator then runs the separate launches (same result as with LOWBIT_K_FUSED=0)."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    k16 = mk(1, 1, 256, 64, "HND", torch.bfloat16, 72).to(cuda_dev)
    assert not L.k_smooth_quant_supported(k16)
    with pytest.raises(NV.LowbitNativeError):
        L.k_smooth_quant(k16)
    assert not NV.lib().lowbit_k_smooth_quant_supported(13000, 64, NV.F16)
    assert NV.lib().lowbit_k_smooth_quant_supported(12800, 64, NV.F16)
    assert not NV.lib().lowbit_k_smooth_quant_supported(6464, 128, NV.F16)
    q = mk(1, 2, 300, 64, "HND", torch.bfloat16, 73).to(cuda_dev)
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q, q, q)
    assert o.dtype == torch.bfloat16 and torch.isfinite(o.float()).all()


# ------------------------------------------------------------------------------------------------ host streaming
@pytest.mark.parametrize("layout,b,hq,hkv,n,d,causal,chunks", [
    ("HND", 2, 4, 4, 700, 64, False, None),
    ("HND", 1, 8, 2, 512, 128, True, 4),     # GQA: q heads follow their kv head
    ("NHD", 3, 4, 4, 300, 64, False, 8),     # NHD: whole batch entries only
    ("HND", 2, 6, 6, 256, 64, True, 1),      # one chunk == the plain call
])
def test_host_streaming_matches_device_call(L, cuda_dev, layout, b, hq, hkv, n, d, causal, chunks):
    """lowbit_fa_host (pinned host tensors, chunked over (batch, head-group) on three streams) is bit-identical
    to one operator call on device tensors."""
    q = mk(b, hq, n, d, layout, torch.float16, 11).pin_memory()
    k = mk(b, hkv, n, d, layout, torch.float16, 12, bias=2.0).pin_memory()
    v = mk(b, hkv, n, d, layout, torch.float16, 13).pin_memory()
    ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                             is_causal=causal)
    out = L.lowbit_fa_host(q, k, v, tensor_layout=layout, chunks=chunks, is_causal=causal, device=cuda_dev)
    torch.cuda.synchronize()
    assert out.device.type == "cpu" and out.shape == q.shape
    assert torch.equal(out, ref.cpu())
    # another operator through the same entry point
    out4 = L.lowbit_fa_host(q, k, v, op=L.lowbit_fa_q_int8_k_int4_pv_fp16, tensor_layout=layout, is_causal=causal,
                            device=cuda_dev)
    ref4 = L.lowbit_fa_q_int8_k_int4_pv_fp16(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                             is_causal=causal)
    torch.cuda.synchronize()
    assert torch.equal(out4, ref4.cpu())
    # return_lse: (out, lse [B, Hq, N] fp32 on the host), eagerly and as a replayed graph -- bit-identical to the device call
    ref_o, ref_lse = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev),
                                                         tensor_layout=layout, is_causal=causal, return_lse=True)
    lse_keep = None
    for graph in (False, True, True, True):
        o2, lse2 = L.lowbit_fa_host(q, k, v, tensor_layout=layout, chunks=chunks, is_causal=causal, device=cuda_dev,
                                    return_lse=True, graph=graph, out=out4, lse_out=lse_keep)
        torch.cuda.synchronize()
        lse_keep = lse2
        assert lse2.device.type == "cpu" and tuple(lse2.shape) == (b, hq, n)
        assert torch.equal(o2, ref_o.cpu()) and torch.equal(lse2, ref_lse.cpu())


@pytest.mark.parametrize("layout,b,hq,hkv,n,d,causal,chunks", [
    ("HND", 2, 4, 4, 700, 64, False, 8),
    ("HND", 1, 8, 2, 512, 128, True, 4),
    ("NHD", 3, 4, 4, 300, 64, False, None),
])
def test_host_streaming_replays_a_captured_graph(L, cuda_dev, layout, b, hq, hkv, n, d, causal, chunks):
    """graph=True: a caller that returns with the same pinned buffers gets the pipeline as one CUDA graph (captured at
    the second call, replayed afterwards).  The buffers are refilled in place between calls: every replay must read the
    new contents and stay bit-identical to the plain operator call; the default (graph=False) never captures."""
    from lowbit_quant_fa2_paddle_b200 import host as H
    H.drop_graphs()
    q = mk(b, hq, n, d, layout, torch.float16, 31).pin_memory()
    k = mk(b, hkv, n, d, layout, torch.float16, 32, bias=2.0).pin_memory()
    v = mk(b, hkv, n, d, layout, torch.float16, 33).pin_memory()
    out = torch.empty(q.shape, dtype=q.dtype).pin_memory()
    for it in range(4):
        if it:  # refill in place: same addresses, new values
            q.copy_(mk(b, hq, n, d, layout, torch.float16, 40 + it))
            k.copy_(mk(b, hkv, n, d, layout, torch.float16, 50 + it, bias=1.0))
            v.copy_(mk(b, hkv, n, d, layout, torch.float16, 60 + it))
        out.zero_()
        got = L.lowbit_fa_host(q, k, v, out=out, tensor_layout=layout, chunks=chunks, is_causal=causal, device=cuda_dev,
                               graph=True)
        ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout,
                                                 is_causal=causal)
        torch.cuda.synchronize()
        assert got is out and torch.equal(out, ref.cpu()), f"call {it}"
    entries = [e for e in H._graphs.values()]
    assert len(entries) == 1 and entries[0][1] is not None and entries[0][0] == 4  # seen 4 times, graph captured
    out.zero_()
    L.lowbit_fa_host(q, k, v, out=out, tensor_layout=layout, chunks=chunks, is_causal=causal, device=cuda_dev)
    torch.cuda.synchronize()
    assert torch.equal(out, ref.cpu()) and entries[0][0] == 4
    H.drop_graphs()


# ------------------------------------------------------------------------------------------------ triton_gpu rounding
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bits", [8, 4])
def test_triton_gpu_division_mode_is_within_one_step_of_ieee(L, cuda_dev, dtype, bits):
    """backend="triton_gpu" replaces the two IEEE divisions of Q1 by PTX div.full.f32 (<= 2 ulp), which is what Triton
    emits for the reference kernels on a GPU.  No CPU restatement of div.full exists (it starts from the hardware
    reciprocal table), so the bit-exact check against the JIT-compiled reference lives in tools/ref_on_b200.py; here:
    scales within 2 ulp of the IEEE ones, codes within one step, and only where the quotient sits on a rounding
    boundary (a tiny fraction)."""
    q = mk(2, 3, 1000, 64, "HND", dtype, 21).to(cuda_dev)
    k = mk(2, 3, 1000, 64, "HND", dtype, 22, bias=1.5).to(cuda_dev)
    km = L.k_mean(k)
    fn = L.per_block_int8 if bits == 8 else L.per_block_int4_unpack
    a = fn(q, k, km=km, backend="triton")
    g = fn(q, k, km=km, backend="triton_gpu")
    for i in (1, 3):
        ulp = (a[i].view(torch.int32) - g[i].view(torch.int32)).abs().max().item()
        assert ulp <= 2, f"scale differs by {ulp} ulp"
    for i in (0, 2):
        diff = (a[i].int() - g[i].int()).abs()
        assert diff.max().item() <= 1
        assert diff.float().mean().item() < 1e-3


# ------------------------------------------------------------------------------------------------ fused preparation
# ------------------------------------------------------------------------------------------------ varlen (packed) path
VARLEN = golden_names("varlen_")


@pytest.mark.parametrize("name", VARLEN)
def test_varlen_matches_reference_kernel_golden(L, cuda_dev, name):
    """Packed (cu_seqlens) quantizer: codes, scales (reference block-major layout) and scale offsets bit-exact against
    the reference's varlen kernel; attention over the same codes within the padded operator's tolerance."""
    g = load_golden(name)
    dev = cuda_dev
    q, k, v = (g[n].to(dev) for n in ("q", "k", "v"))
    cu_q, cu_k = g["cu_q"].to(dev), g["cu_k"].to(dev)
    mq = int((g["cu_q"][1:] - g["cu_q"][:-1]).max())
    mk_ = int((g["cu_k"][1:] - g["cu_k"][:-1]).max())
    km = L.k_mean_varlen(k)
    assert torch.equal(km.cpu(), g["km"])
    got = L.per_block_int8_varlen(q, k, cu_q, cu_k, mq, mk_, sm_scale=g["sm_scale"], km=km)
    for t, n in zip(got, ("q_int8", "q_scale", "k_int8", "k_scale", "cu_q_scale", "cu_k_scale")):
        assert torch.equal(t.cpu().to(g[n].dtype), g[n]), n
    causal = bool(g["causal"])
    o = L.forward_varlen(got[0], got[2], v.to(torch.float16), cu_q, cu_k, mq, got[1], got[3], got[4], got[5],
                         output_dtype=g["o"].dtype, causal=causal, compat_tail=not causal)
    tol = 4e-3 if g["o"].dtype == torch.float16 else 3.2e-2
    err = (o.cpu().float() - g["o"].float()).abs().max().item()
    assert err <= tol, f"max-abs {err}"
    assert cos_sim(o.cpu(), g["o"]) >= 0.999


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype,d", [(torch.float16, 64), (torch.bfloat16, 128), (torch.float16, 80)])
def test_varlen_api_vs_oracle_and_sdpa(L, cuda_dev, causal, dtype, d):
    """lowbit_fa_varlen end to end (batch-global K mean, per-sequence blocks, empty and one-token sequences, int64
    cu_seqlens, padded head_dim) against the CPU oracle of core.py:356-491 and per-sequence FP32 SDPA."""
    from oracle import attention as OA
    from oracle import varlen as OV
    lq = [300, 0, 1, 129, 64] + ([] if causal else [7])
    lk = lq if causal else [77, 0, 5, 256, 1, 0]   # last non-causal sequence: queries without keys -> zeros
    hq, hkv = 4, 2
    g = torch.Generator().manual_seed(41)
    q = torch.randn(sum(lq), hq, d, generator=g).to(dtype)
    k = (torch.randn(sum(lk), hkv, d, generator=g) + 1.5 * torch.randn(1, hkv, d, generator=g)).to(dtype)
    v = torch.randn(sum(lk), hkv, d, generator=g).to(dtype)
    cu_q = [0] + torch.tensor(lq).cumsum(0).tolist()
    cu_k = [0] + torch.tensor(lk).cumsum(0).tolist()
    o = L.lowbit_fa_varlen(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev),
                           torch.tensor(cu_q, dtype=torch.int64, device=cuda_dev),
                           torch.tensor(cu_k, dtype=torch.int32, device=cuda_dev), max(lq), max(lk), is_causal=causal)
    assert o.shape == q.shape and o.dtype == dtype
    ref = OV.lowbit_fa_varlen_api(q, k, v, cu_q, cu_k, causal, compat_tail=False, pv_accum="fp32")
    tol = 4e-3 if dtype == torch.float16 else 3.2e-2
    assert (o.cpu().float() - ref.float()).abs().max().item() <= tol
    for i in range(len(lq)):
        a, b, c, e = cu_q[i], cu_q[i + 1], cu_k[i], cu_k[i + 1]
        if b == a or e == c:
            continue
        sd = OA.sdpa_fp32(q[a:b].unsqueeze(0), k[c:e].unsqueeze(0), v[c:e].unsqueeze(0), "NHD", causal)
        assert cos_sim(o[a:b].cpu(), sd[0]) >= 0.999


# ------------------------------------------------------------------------------------------------ dispatcher / CUDA-API names
def test_sageattn_dispatcher_and_cuda_api_names(L, cuda_dev):
    """sageattn / lowbit_fa_attn (core.py:82-190) and the *_fp16_cuda signature (core.py:495-731) run the sm_100a kernel:
    identical to the _triton entry point under the CUDA quantizer's rounding, SDPA keywords ignored, lse returned."""
    q = mk(2, 4, 300, 64, "HND", torch.float16, 51).to(cuda_dev)
    k = mk(2, 2, 300, 64, "HND", torch.float16, 52, bias=2.0).to(cuda_dev)
    v = mk(2, 2, 300, 64, "HND", torch.float16, 53).to(cuda_dev)
    ref, lse_ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v, quantization_backend="cuda", is_causal=True,
                                                      return_lse=True)
    o, lse = L.sageattn(q, k, v, is_causal=True, return_lse=True, attn_mask=None, dropout_p=0.0)
    assert torch.equal(o, ref) and torch.equal(lse, lse_ref)
    for acc in ("fp16", "fp16+fp32", "fp32"):
        o2 = L.lowbit_fa_qk_int8_pv_fp16_cuda(q, k, v, is_causal=True, pv_accum_dtype=acc, qk_quant_gran="per_warp",
                                              smooth_v=True)
        assert torch.equal(o2, ref)
    from oracle import attention as OA
    sd = OA.sdpa_fp32(q.cpu(), k.cpu(), v.cpu(), "HND", True)
    assert cos_sim(o.cpu(), sd) >= 0.999


# ------------------------------------------------------------------------------------------------ dynamic K bit allocation
def _mixed_k(b, h, n, d, layout, dtype, seed):
    """K whose 64-row blocks fall into all three width classes after smoothing."""
    k = mk(b, h, n, d, "HND", torch.float32, seed, bias=2.0)
    km = k.mean(dim=2, keepdim=True)
    g = torch.Generator().manual_seed(seed + 1)
    nblk = (n + 63) // 64
    w = torch.tensor([0.1, 1.0, 3.0])[torch.randint(0, 3, (b, h, nblk), generator=g)]
    w = w.repeat_interleave(64, dim=2)[:, :, :n].unsqueeze(-1)
    k = (km + (k - km) * w).to(dtype)
    return k if layout == "HND" else k.permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("layout", ["HND", "NHD"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,d", [(512, 64), (333, 128), (64, 64), (1, 128)])
def test_mixed_k_quantizer_bit_exact(L, cuda_dev, layout, dtype, n, d):
    """lowbit_quant_k_mixed: per-block widths, scales and the container bytes (row prefixes in the kernel's expansion
    order) are bit-exact against oracle.quant.quant_k_mixed + pack_mixed, with thresholds and with an imposed map."""
    from oracle import quant as OQ
    k = _mixed_k(2, 3, n, d, layout, dtype, 61)
    km_ref = OQ.k_mean(k, layout)
    km = L.k_mean(k.to(cuda_dev), layout)
    assert torch.equal(km.cpu(), km_ref)
    codes, scale, kb = L.per_block_k_mixed(k.to(cuda_dev), km, tensor_layout=layout)
    rc, rs, rb = OQ.quant_k_mixed(k, km_ref, None, 64, layout)
    assert torch.equal(kb.cpu(), rb) and torch.equal(scale.cpu(), rs)
    assert torch.equal(codes.cpu(), OQ.pack_mixed(rc, rb, 64, layout))
    assert torch.equal(OQ.unpack_mixed(codes.cpu(), rb, 64, layout), rc)
    if n >= 64:
        assert set(rb.unique().tolist()) == {2, 4, 8}
    imposed = torch.tensor([2, 4, 8], dtype=torch.int32)[torch.randint(0, 3, rb.shape)]
    codes, scale, kb = L.per_block_k_mixed(k.to(cuda_dev), km, kbits=imposed, tensor_layout=layout)
    rc, rs, rb = OQ.quant_k_mixed(k, km_ref, imposed, 64, layout)
    assert torch.equal(kb.cpu(), imposed) and torch.equal(scale.cpu(), rs)
    assert torch.equal(codes.cpu(), OQ.pack_mixed(rc, rb, 64, layout))


@pytest.mark.parametrize("layout,hq,hkv,n,d,causal,pv", [
    ("HND", 2, 2, 512, 64, False, "fp16"),
    ("HND", 4, 2, 384, 128, True, "fp16"),
    ("NHD", 2, 2, 200, 64, False, "fp16"),    # ragged tail
    ("NHD", 2, 1, 333, 128, True, "fp8"),
    ("HND", 2, 2, 1024, 64, True, "fp8"),
])
def test_mixed_k_attention_identical_to_int8_path(L, cuda_dev, layout, hq, hkv, n, d, causal, pv):
    """The mixed-width kernel path (TMA box per bit width, 8/4/2-bit expansion in shared memory, power-of-two factor
    folded into the scale) yields exactly the same integer scores as the INT8 path fed the unpacked codes; the
    effective scale differs by that power of two, and with it the rounding of the exponent offset that carries the
    score bias (see test_packed_int4_k_matches_unpacked): O within 2 fp16 ulps of its magnitude (fp16 P; one e4m3
    step of P for FP8 P.V), lse within 1e-3."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    from oracle import quant as OQ
    q = mk(1, hq, n, d, layout, torch.float16, 71).to(cuda_dev)
    k = _mixed_k(1, hkv, n, d, layout, torch.float16, 72).to(cuda_dev)
    v = mk(1, hkv, n, d, layout, torch.float16, 73).to(cuda_dev)
    km = L.k_mean(k, layout)
    qc, qs, _, _ = L.per_block_int8(q, k, km=km, tensor_layout=layout)
    kc, ks, kb = L.per_block_k_mixed(k, km, tensor_layout=layout)
    k_unp = OQ.unpack_mixed(kc.cpu(), kb.cpu(), 64, layout).to(cuda_dev)
    kw = {}
    if pv == "fp8":
        v, vs, _ = L.per_channel_fp8(v, layout, smooth_v=False)
        kw = dict(pv_mode=NV.PV_E4M3, v_scale=vs)
    fn = L.forward_causal if causal else L.forward
    o_mix, lse_mix = fn(qc, kc, v, qs, ks, tensor_layout=layout, return_lse=True, qk_mode=NV.QK_Q8KMIX, kbits=kb, **kw)
    # the mixed-width path runs 32-key steps: compare with the same kernel family (narrow=True)
    o_i8, lse_i8 = fn(qc, k_unp, v, qs, ks, tensor_layout=layout, return_lse=True, narrow=True, **kw)
    omax = max(o_i8.float().abs().max().item(), 2.0 ** -10)
    assert (o_mix.float() - o_i8.float()).abs().max().item() <= (2.0 * 2.0 ** -10 if pv == "fp16" else 2.0 ** -4) * omax
    assert (lse_mix - lse_i8).abs().max().item() <= (1e-3 if pv == "fp16" else 3e-2)
    if d == 64:  # the default 64-key-step kernel: same softmax, different reference maxima / exp2 pipe
        o_w, lse_w = fn(qc, k_unp, v, qs, ks, tensor_layout=layout, return_lse=True, **kw)
        assert (o_w.float() - o_i8.float()).abs().max().item() <= (2e-3 if pv == "fp16" else 5e-2)
        assert (lse_w - lse_i8).abs().max().item() <= (1e-3 if pv == "fp16" else 3e-2)


@pytest.mark.parametrize("d,causal,pv", [(64, False, "fp16"), (128, True, "fp16"), (128, True, "fp8")])
def test_dynamic_k_api_vs_oracle_and_sdpa(L, cuda_dev, d, causal, pv):
    from oracle import attention as OA
    q = mk(1, 4, 600, d, "HND", torch.float16, 81)
    k = _mixed_k(1, 2, 600, d, "HND", torch.float16, 82)
    v = mk(1, 2, 600, d, "HND", torch.float16, 83)
    o, lse = L.lowbit_fa_q_int8_k_dynamic(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), is_causal=causal,
                                          return_lse=True, pv=pv)
    ref, lse_ref = OA.lowbit_fa_api(q, k, v, "HND", causal, return_lse=True, compat_tail=False, pv_accum="fp32",
                                    qk="mixed", pv=pv)
    tol = 4e-3 if pv == "fp16" else 0.125 * float(v.abs().max())
    assert (o.cpu().float() - ref.float()).abs().max().item() <= tol
    assert (lse.cpu() - lse_ref).abs().max() <= (2e-3 if pv == "fp16" else 3e-2)
    # INT2 / INT4 blocks are a coarse approximation by design: accuracy against exact attention is reported, with a
    # loose floor, not asserted at the INT8 level
    sd = OA.sdpa_fp32(q, k, v, "HND", causal)
    assert cos_sim(o.cpu(), sd) >= 0.98


# ------------------------------------------------------------------------------------------------ kernel selection
@pytest.mark.parametrize("env", [{"LOWBIT_ATTN_N64": "0"}, {"LOWBIT_ATTN_PF": "0"}, {"LOWBIT_ATTN_PF": "3"}])
def test_attention_suite_under_every_kernel_selection(cuda_dev, env):
    """head_dim 64 has two kernels and the 64-key-step one has a tunable share of FMA-pipe exp2: LOWBIT_ATTN_N64=0
    (32-key steps everywhere), LOWBIT_ATTN_PF=0 / 3 (none / 3 of every 8 score pairs off the MUFU pipe).  The switches
    are read once per process, so the attention tests are re-run in a child process under each setting."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", "golden or api_vs_oracle or fp8_pv or packed_int4 or ring_partial or tiny_sequences "
                              "or tail_masking or large_magnitude",
                        "--deselect", "tests/test_gpu_parity.py::test_attention_suite_under_every_kernel_selection"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n", [1, 31, 64, 65, 129])
@pytest.mark.parametrize("d", [64, 128])
def test_tiny_sequences_every_k_format(L, cuda_dev, n, d):
    """One to a few key blocks (prologue-only pipelines: expander warp, K stage ring, masked tail) for the packed INT4,
    mixed-width and FP8-PV operators, causal and not, against the CPU oracle."""
    from oracle import attention as OA
    q = mk(1, 2, n, d, "HND", torch.float16, 101)
    k = mk(1, 2, n, d, "HND", torch.float16, 102, bias=1.0)
    v = mk(1, 2, n, d, "HND", torch.float16, 103)
    dq, dk, dv = q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev)
    for causal in (False, True):
        o = L.lowbit_fa_qk_int4_pv_fp16_triton(dq, dk, dv, is_causal=causal)
        ref = OA.lowbit_fa_api(q, k, v, "HND", causal, compat_tail=False, pv_accum="fp32", qk="int4")
        assert (o.cpu().float() - ref.float()).abs().max().item() <= 4e-3
        o = L.lowbit_fa_q_int8_k_dynamic(dq, dk, dv, is_causal=causal)
        ref = OA.lowbit_fa_api(q, k, v, "HND", causal, compat_tail=False, pv_accum="fp32", qk="mixed")
        assert (o.cpu().float() - ref.float()).abs().max().item() <= 4e-3
        o = L.lowbit_fa_qk_int4_pv_fp8(dq, dk, dv, is_causal=causal)
        ref = OA.lowbit_fa_api(q, k, v, "HND", causal, compat_tail=False, pv_accum="fp32", qk="int4", pv="fp8")
        assert (o.cpu().float() - ref.float()).abs().max().item() <= 0.125 * float(v.abs().max())


# ------------------------------------------------------------------------------------------------ SDPA plug-in (f-3)
@pytest.mark.parametrize("op_name", ["lowbit_fa_qk_int8_pv_fp16_triton", "lowbit_fa_q_int8_k_int4_pv_fp16"])
def test_sdpa_monkey_patch_cogvideox_shape(L, cuda_dev, op_name):
    """The reference's plug-in path (example/sageattn_cogvideo.py:9-14): F.scaled_dot_product_attention replaced by the
    operator and called the way diffusers calls it -- positional q, k, v in [B,H,N,D], bf16, N = 17776 (not a multiple
    of 64 or 128), with attn_mask / dropout_p / is_causal keywords.  Checked against torch's own SDPA in fp32 on a head
    slice (max-abs, MSE, cos-sim) and at the block level through examples/cogvideox_block.Attention."""
    import torch.nn.functional as F
    from lowbit_quant_fa2_paddle_b200.plugin import patch_sdpa
    torch.manual_seed(0)
    b, h, n, d = 1, 6, 17776, 64
    q, k, v = (torch.randn(b, h, n, d, device=cuda_dev, dtype=torch.bfloat16) for _ in range(3))
    k = k + 0.5 * torch.randn(1, h, 1, d, device=cuda_dev, dtype=torch.bfloat16)  # a channel bias for the smoothing
    orig = F.scaled_dot_product_attention
    with patch_sdpa(getattr(L, op_name)):
        assert F.scaled_dot_product_attention is not orig
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    assert F.scaled_dot_product_attention is orig, "the patch must be undone"
    assert o.shape == q.shape and o.dtype == torch.bfloat16
    ref = orig(q[:, :2].float(), k[:, :2].float(), v[:, :2].float())
    got = o[:, :2].float()
    cos = cos_sim(got.cpu(), ref.cpu())
    assert cos >= (0.999 if "int8_pv" in op_name else 0.98), cos  # 4-bit K codes: format error, not kernel error
    assert (got - ref).pow(2).mean().item() <= (2e-5 if "int8_pv" in op_name else 1e-3)
    # block level: the synthetic CogVideoX attention block, patched vs not
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("cogvideox_block", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "examples", "cogvideox_block.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    blk = mod.Attention().to(cuda_dev, torch.bfloat16).eval()
    x = torch.randn(1, 4096 + 48, 3072, device=cuda_dev, dtype=torch.bfloat16)
    with torch.no_grad():
        y0 = blk(x)
        with patch_sdpa(getattr(L, op_name)):
            y1 = blk(x)
    assert cos_sim(y1.float().cpu(), y0.float().cpu()) >= 0.9999


def test_misaligned_views_and_wide_heads(L, cuda_dev):
    """A view whose storage offset is not a multiple of 16 bytes (a column slice of a fused buffer: legal for the
    reference's Triton kernels) goes through the operator (it is re-aligned by a copy), the raw quantizer entry point
    refuses it with a message instead of faulting, and compute_scale rejects head_dim > 128 like the operators do."""
    from lowbit_quant_fa2_paddle_b200 import _native as NV
    torch.manual_seed(11)
    fused = torch.randn(1, 2, 256, 3 * 64 + 4, dtype=torch.float16, device=cuda_dev)
    q, k, v = fused[..., 4:68], fused[..., 68:132], fused[..., 132:196]
    assert q.data_ptr() % 16 != 0
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q, k, v)
    ref = L.lowbit_fa_qk_int8_pv_fp16_triton(q.contiguous(), k.contiguous(), v.contiguous())
    assert torch.equal(o, ref)
    with pytest.raises(NV.LowbitNativeError, match="16-byte"):
        L.k_mean(k)
    with pytest.raises(ValueError):
        L.compute_scale(torch.randn(1, 1, 8, 160, dtype=torch.float16, device=cuda_dev))


# ------------------------------------------------------------------------------------------------ E4: the FP16 class
@pytest.mark.parametrize("case", [
    # b, hq, hkv, n, d, layout, causal, dtype
    (1, 2, 2, 512, 64, "HND", False, torch.float16),
    (2, 4, 2, 333, 128, "NHD", True, torch.float16),
    (1, 2, 2, 200, 64, "NHD", True, torch.bfloat16),
    (1, 3, 3, 1030, 128, "HND", False, torch.bfloat16),
    (1, 2, 1, 65, 80, "HND", False, torch.float16),   # head_dim padded to 128
    (1, 1, 1, 1, 64, "HND", True, torch.float16),
])
def test_fp16_class_attention_vs_sdpa(L, cuda_dev, case):
    """lowbit_fa_fp16 (QK_F16: tcgen05 kind::f16 Q.K^T over the fp16 / bf16 inputs, fp32 scores) -- what
    lowbit_fa_multi_precision runs for the "FP16" class, where the reference calls plain SDPA (core.py:1075-1076).
    Against fp32 SDPA: max-abs <= 2e-3 (fp16) / 1.6e-2 (bf16 output rounding), cos >= 0.9999; lse <= 2e-3."""
    from oracle import attention as OA
    b, hq, hkv, n, d, layout, causal, dtype = case
    q, k, v = mk(b, hq, n, d, layout, dtype, 91), mk(b, hkv, n, d, layout, dtype, 92, bias=1.0), mk(b, hkv, n, d, layout, dtype, 93)
    o, lse = L.lowbit_fa_fp16(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), tensor_layout=layout, is_causal=causal,
                              return_lse=True)
    ref, lref = OA.sdpa_fp32(q, k, v, layout, causal, return_lse=True)
    assert o.shape == q.shape and o.dtype == dtype
    assert (o.cpu().float() - ref.float()).abs().max().item() <= (2e-3 if dtype == torch.float16 else 1.6e-2)
    assert (lse.cpu() - lref).abs().max().item() <= 2e-3
    if n >= 64:
        assert cos_sim(o.cpu(), ref) >= 0.9999


def test_multi_precision_routes_every_class(L, cuda_dev):
    """select_quantization's three classes (core.py:1050-1061) each reach their own operator: inputs scaled so that the
    mean of max|x|/127 lands above 0.2 (FP16: un-quantized kernel, bit-identical to lowbit_fa_fp16), between (INT8),
    and below 0.05 (INT4)."""
    base = [mk(1, 2, 256, 64, "HND", torch.float16, s).to(cuda_dev) for s in (31, 32, 33)]
    for scale, kind, fn in ((12.0, "FP16", L.lowbit_fa_fp16), (2.5, "INT8", L.lowbit_fa_qk_int8_pv_fp16_triton),
                            (0.5, "INT4", L.lowbit_fa_qk_int4_pv_fp16_triton)):
        q, k, v = ((t.float() * scale).half() for t in base)
        assert L.select_quantization(q, k, v) == kind
        assert torch.equal(L.lowbit_fa_multi_precision(q, k, v, sm_scale=0.05), fn(q, k, v, sm_scale=0.05))


# ------------------------------------------------------------------------------------------------ one-call operator
@pytest.mark.gpu
@pytest.mark.parametrize("case", [
    # B, Hq, Hkv, Nq, Nk, D, layout, dtype, causal, entry, lse
    (1, 2, 2, 512, 512, 64, "HND", torch.float16, False, "int8", False),
    (2, 4, 2, 200, 333, 64, "NHD", torch.float16, False, "int8", True),
    (1, 4, 4, 1030, 1030, 64, "HND", torch.bfloat16, True, "int8", True),
    (1, 2, 1, 384, 384, 128, "NHD", torch.float16, True, "int4", False),
    (2, 6, 3, 2200, 2200, 64, "NHD", torch.float16, False, "q8k4", True),
    (1, 32, 32, 4096, 4096, 64, "HND", torch.float16, False, "int8", False),   # large enough for the side-stream fork
    (1, 8, 8, 1024, 1024, 128, "HND", torch.bfloat16, False, "q8k4", True),
], ids=lambda c: f"{c[9]}-{c[6]}-d{c[5]}-n{c[3]}x{c[4]}-{'c' if c[8] else 'nc'}-{str(c[7]).split('.')[-1]}")
def test_one_call_operator_is_bit_identical_to_the_five_call_path(case, monkeypatch):
    """lowbit_fa_fwd (csrc/op.cu) launches the kernels of the separate entry points with the same arguments: o and lse
    must be bit for bit what the five-call host path gives (LOWBIT_ONE_CALL=0)."""
    import lowbit_quant_fa2_paddle_b200 as L
    from lowbit_quant_fa2_paddle_b200 import core
    B, Hq, Hkv, Nq, Nk, D, layout, dt, causal, entry, want_lse = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    shp = lambda h, n: (B, h, n, D) if layout == "HND" else (B, n, h, D)
    q, k, v = (torch.randn(shp(h, n), generator=g).to(dt).to(dev) for h, n in ((Hq, Nq), (Hkv, Nk), (Hkv, Nk)))
    fn = {"int8": L.lowbit_fa_qk_int8_pv_fp16_triton, "int4": L.lowbit_fa_qk_int4_pv_fp16_triton,
          "q8k4": L.lowbit_fa_q_int8_k_int4_pv_fp16}[entry]
    kw = dict(tensor_layout=layout, is_causal=causal, return_lse=want_lse)
    monkeypatch.setattr(core, "_ONE_CALL", True)
    a = fn(q, k, v, **kw)
    monkeypatch.setattr(core, "_ONE_CALL", False)
    b = fn(q, k, v, **kw)
    torch.cuda.synchronize()
    if want_lse:
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    else:
        assert torch.equal(a, b)
