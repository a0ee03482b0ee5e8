"""CPU: pin the oracle (oracle/*.py) against golden vectors produced by the reference's own Triton
kernels (tools/make_golden.py).  Quantizer codes and scales: bit-exact.  Attention: tolerance below."""
import pytest
import torch

from conftest import cos_sim, golden_names, load_golden
from oracle import attention as A
from oracle import quant as Q

ATTN = golden_names("attn_")


@pytest.mark.parametrize("name", ATTN)
def test_q1_q4_bit_exact(name):
    g = load_golden(name)
    qi, qs, ki, ks = Q.per_block_int8_q1(g["q"], g["k"], g["km"], sm_scale=g["sm_scale"], tensor_layout=g["layout"])
    assert torch.equal(qi, g["q_int8"]) and torch.equal(ki, g["k_int8"])
    assert torch.equal(qs, g["q_scale"]) and torch.equal(ks, g["k_scale"])
    _, _, k4, k4s = Q.per_block_int8_q1(g["q"], g["k"], g["km"], sm_scale=g["sm_scale"], tensor_layout=g["layout"], kbits=4)
    assert torch.equal(k4, g["k_int4"]) and torch.equal(k4s, g["k_int4_scale"])
    assert int(k4.abs().max()) <= 7


@pytest.mark.parametrize("name", [n for n in ATTN if "causal" not in n])
@pytest.mark.parametrize("bits", [8, 4])
def test_q3_per_thread_bit_exact(name, bits):
    g = load_golden(name)
    qi, qs, ki, ks = Q.per_thread(g["q"], g["k"], g["km"], tensor_layout=g["layout"], bits=bits)
    p = f"pt{bits}_"
    assert torch.equal(qi, g[p + "q"]) and torch.equal(ki, g[p + "k"])
    assert torch.equal(qs, g[p + "qs"]) and torch.equal(ks, g[p + "ks"])


@pytest.mark.parametrize("name", ATTN)
def test_attention_emulator_matches_reference_kernel(name):
    """Tolerance: the emulator restates the block algorithm in fp32 torch ops; the interpreter's exp2 and
    fp16 dot differ in the last bit -> allow 4 output ulps (fp16: 2^-10 relative at |o|<=4; bf16: 2^-7)."""
    g = load_golden(name)
    o, lse2 = A.attn_block_emulator(g["q_int8"], g["k_int8"], g["v"].to(torch.float16), g["q_scale"], g["k_scale"],
                                    g["layout"], bool(g["causal"]), output_dtype=g["o"].dtype, return_lse=True)
    tol = 4e-3 if g["o"].dtype == torch.float16 else 3.2e-2
    assert (o.float() - g["o"].float()).abs().max() <= tol
    assert (lse2 - g["lse2"]).abs().max() <= 1e-4
    assert cos_sim(o, g["o"]) > 0.99999
    o4, _ = A.attn_block_emulator(g["q_int8"], g["k_int4"], g["v"].to(torch.float16), g["q_scale"], g["k_int4_scale"],
                                  g["layout"], bool(g["causal"]), output_dtype=g["o"].dtype)
    assert (o4.float() - g["o_k4"].float()).abs().max() <= tol


@pytest.mark.parametrize("name", ATTN)
def test_api_glue_vs_sdpa(name):
    """End-to-end oracle (core.py:269-352 restated) stays close to FP32 SDPA -- the reference's own
    acceptance procedure (example/test_sageattn_operator.py:92-94) with an explicit threshold."""
    g = load_golden(name)
    causal = bool(g["causal"])
    o, lse = A.lowbit_fa_api(g["q"], g["k"], g["v"], g["layout"], causal, return_lse=True, compat_tail=False)
    ref, lse_ref = A.sdpa_fp32(g["q"], g["k"], g["v"], g["layout"], causal, return_lse=True)
    assert cos_sim(o, ref) > 0.999
    assert (lse - lse_ref).abs().max() < 0.05


@pytest.mark.parametrize("bit", [2, 4, 8])
def test_q5_kivi_pack_bit_exact(bit):
    g = load_golden(f"kivi_b{bit}")
    code, scale, mn = Q.kivi_quantize_and_pack(g["data"], int(g["group_size"]), bit)
    assert torch.equal(code, g["code"])
    assert torch.equal(scale.view(torch.int16), g["scale"].view(torch.int16))
    assert torch.equal(mn.view(torch.int16), g["mn"].view(torch.int16))
    x = Q.kivi_unpack_and_dequant(code, scale, mn, 32, bit)
    err = (x.float() - g["data"].float()).abs().max()
    assert err <= (g["data"].float().amax() - g["data"].float().amin()) / (2 ** bit - 1)


def test_pack_unpack_roundtrip():
    for bits, lim in ((4, 7), (2, 1)):
        c = torch.randint(-lim, lim + 1, (2, 3, 17, 64), dtype=torch.int8)
        p = Q.pack_codes(c, bits)
        assert p.shape[-1] == 64 * bits // 8
        assert torch.equal(Q.unpack_codes(p, bits), c)


def test_k_mean_order_independent():
    torch.manual_seed(1)
    k = (torch.randn(1, 2, 300, 64) * 3 + 5).half()
    km = Q.k_mean(k)
    perm = torch.randperm(300)
    assert torch.equal(km, Q.k_mean(k[:, :, perm]))
    assert (km.float() - k.float().mean(dim=2, keepdim=True)).abs().max() < 4e-3
    kn = k.permute(0, 2, 1, 3).contiguous()
    assert torch.equal(Q.k_mean(kn, "NHD").permute(0, 2, 1, 3), km)


# ------------------------------------------------------------------------------------------------ varlen (packed) path
VARLEN = golden_names("varlen_")


@pytest.mark.parametrize("name", VARLEN)
def test_varlen_oracle_matches_reference_kernels(name):
    """oracle/varlen.py against the reference's own varlen quantizer + attention kernels (interpreter):
    codes / scales / scale offsets bit-exact, attention within 4 output ulps."""
    from oracle import varlen as OV
    g = load_golden(name)
    cu_q, cu_k = g["cu_q"].tolist(), g["cu_k"].tolist()
    km = OV.k_mean_varlen(g["k"])
    assert torch.equal(km, g["km"])
    qi, qs, ki, ks, cqs, cks = OV.per_block_int8_varlen(g["q"], g["k"] - km, cu_q, cu_k, sm_scale=g["sm_scale"])
    assert torch.equal(qi, g["q_int8"]) and torch.equal(ki, g["k_int8"])
    assert torch.equal(qs, g["q_scale"]) and torch.equal(ks, g["k_scale"])
    assert cqs == g["cu_q_scale"].tolist() and cks == g["cu_k_scale"].tolist()
    o = OV.attn_varlen(qi, ki, g["v"].to(torch.float16), cu_q, cu_k, qs, ks, cqs, cks, bool(g["causal"]), g["o"].dtype)
    tol = 4e-3 if g["o"].dtype == torch.float16 else 3.2e-2
    assert (o.float() - g["o"].float()).abs().max() <= tol
    assert cos_sim(o, g["o"]) > 0.99999


@pytest.mark.parametrize("name", VARLEN)
def test_varlen_api_glue_vs_sdpa(name):
    """core.py:356-491 restated end to end stays close to per-sequence FP32 SDPA."""
    from oracle import varlen as OV
    g = load_golden(name)
    cu_q, cu_k = g["cu_q"].tolist(), g["cu_k"].tolist()
    causal = bool(g["causal"])
    o = OV.lowbit_fa_varlen_api(g["q"], g["k"], g["v"], cu_q, cu_k, causal, compat_tail=False, pv_accum="fp32")
    for i in range(len(cu_q) - 1):
        a, b, c, e = cu_q[i], cu_q[i + 1], cu_k[i], cu_k[i + 1]
        ref = A.sdpa_fp32(g["q"][a:b].unsqueeze(0), g["k"][c:e].unsqueeze(0), g["v"][c:e].unsqueeze(0), "NHD", causal)
        assert cos_sim(o[a:b], ref[0]) > 0.999


def test_mixed_container_pack_unpack_roundtrip():
    """oracle.quant.pack_mixed / unpack_mixed (the D-byte mixed-width K container of lowbit_quant_k_mixed) are inverse
    for every width class, both layouts and head dims, and a block of width w only touches the first D*w/8 bytes."""
    g = torch.Generator().manual_seed(5)
    for layout in ("HND", "NHD"):
        for d in (64, 128):
            n = 200
            kb = torch.tensor([2, 4, 8], dtype=torch.int32)[torch.randint(0, 3, (1, 2, (n + 63) // 64), generator=g)]
            lim = torch.tensor([0, 0, 1, 0, 7, 0, 0, 0, 127])[kb.long()].repeat_interleave(64, dim=2)[:, :, :n]
            c = (torch.rand(1, 2, n, d, generator=g) * 2 - 1) * lim[..., None]
            c = c.round().to(torch.int8)
            if layout == "NHD":
                c = c.permute(0, 2, 1, 3).contiguous()
            p = Q.pack_mixed(c, kb, 64, layout)
            assert torch.equal(Q.unpack_mixed(p, kb, 64, layout), c)
            ph = p if layout == "HND" else p.permute(0, 2, 1, 3)
            rows = kb.repeat_interleave(64, dim=2)[:, :, :n]
            for w in (2, 4):
                tail = ph[..., d * w // 8:][rows == w]
                assert int(tail.abs().sum()) == 0
