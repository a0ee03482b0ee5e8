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
    """Q6: e4m3 bytes, scales and means bit-exact against the IEEE restatement of fused.cu (parity unpinned:
    the reference CUDA cannot be built here)."""
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
