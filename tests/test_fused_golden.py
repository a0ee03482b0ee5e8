"""Rows Q2 / Q6 pinned to the reference's OWN CUDA kernels.

tests/golden/fused_*.npz hold inputs and outputs of csrc/fused/fused.cu (QuantInt8Kernel :64-198, SubMeanKernel
:200-261, TransposePadPermuteKernel :263-330, MeanScaleKernel :332-428), compiled unmodified from /root/reference by
oracle/build_ref_fused.py (plain nvcc -O3: IEEE arithmetic) and run on a B200 by tools/make_golden_fused.py exactly as
src/quant.py calls them (per_block_int8 :70-98, per_warp_int8 :147-172, sub_mean :203-208, per_channel_fp8 :254-291).

  not gpu:  the CPU oracle (oracle/quant.py) against the goldens   -- pins the restatement
  gpu:      this repository's CUDA kernels against the goldens     -- parity proper, through the C ABI

Codes and scales: bit-exact.  The smooth_v mean is an fp32 block reduction whose order the reference does not fix
(blockReduceSum of fp32 partials; ours is an exact sum rounded once): compared to 2e-6 relative to the largest mean, the
scale to 2 ulp, and the codes that follow from it to one e4m3 step on at most 0.1 % of positions.
"""
import glob
import os

import numpy as np
import pytest
import torch

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fused_*.npz")))
NAMES = [os.path.basename(p)[len("fused_"):-len(".npz")] for p in GOLD]


def load(path):
    z = np.load(path, allow_pickle=False)
    dt = {"float16": torch.float16, "bfloat16": torch.bfloat16}[str(z["dtype"])]
    layout = str(z["layout"])

    def t16(name):
        return torch.from_numpy(z[name].copy()).view(dt)
    g = {k: torch.from_numpy(z[k].copy()) for k in z.files if k not in ("layout", "dtype", "sm_scale")}
    for k in ("q", "k", "v", "km", "sm_vm", "f8_vt"):
        g[k] = t16(k)
    g["sm_v"] = torch.from_numpy(z["sm_v"].copy()).view(torch.float16)
    for k in ("f8_v8", "f8s_v8"):
        g[k] = torch.from_numpy(z[k].copy()).view(torch.float8_e4m3fn)
    return g, layout, dt, float(z["sm_scale"])


def ulp_diff(a, b):
    ia, ib = a.contiguous().view(torch.int32).long(), b.contiguous().view(torch.int32).long()
    return (ia - ib).abs().max().item()


def check_per_block(got, g, prefix_q, prefix_k):
    qi, qs, ki, ks = got
    assert torch.equal(qi.cpu(), g[f"{prefix_q}_q_int8"]), "Q codes differ from the reference kernel"
    assert torch.equal(qs.cpu(), g[f"{prefix_q}_q_scale"]), "Q scales differ from the reference kernel"
    assert torch.equal(ki.cpu(), g[f"{prefix_k}_k_int8"]), "K codes differ from the reference kernel"
    assert torch.equal(ks.cpu(), g[f"{prefix_k}_k_scale"]), "K scales differ from the reference kernel"


def check_fp8(v8, vs, vm, g, tag):
    assert torch.equal(vs.cpu(), g[f"{tag}_scale"]) or ulp_diff(vs.cpu(), g[f"{tag}_scale"]) <= (2 if tag == "f8s" else 0)
    a, b = v8.cpu().view(torch.uint8), g[f"{tag}_v8"].view(torch.uint8)
    if tag == "f8":
        assert torch.equal(a, b), "e4m3 V codes differ from the reference kernel"
    else:
        ref = g["f8s_vm"]
        assert (vm.cpu() - ref).abs().max().item() <= 2e-6 * max(1.0, float(ref.abs().max())), \
            "v mean differs from the reference kernel by more than fp32 summation-order noise"
        diff = (a != b)
        assert diff.float().mean().item() <= 1e-3
        assert (a.int() - b.int()).abs().max().item() <= 1, "a code differs by more than one e4m3 step"


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
@pytest.mark.parametrize("path", GOLD, ids=NAMES)
def test_oracle_q2_q6_matches_reference_cuda_kernels(path):
    from oracle import quant as OQ
    g, layout, dt, sm = load(path)
    q, k, v, km = g["q"], g["k"], g["v"], g["km"]
    check_per_block(OQ.per_block_int8_q2(q, k, km, 128, 64, sm, layout), g, "pb", "pb")
    got = OQ.per_block_int8_q2(q, k, None, 128, 64, sm, layout)
    assert torch.equal(got[2], g["pbn_k_int8"]) and torch.equal(got[3], g["pbn_k_scale"])
    check_per_block(OQ.per_warp_int8_q2(q, k, km, layout), g, "pw", "pw")
    vs, _ = OQ.sub_mean_given(v, g["sm_vm"], layout)
    assert torch.equal(vs, g["sm_v"]), "sub_mean differs from the reference kernel"
    for smooth, tag in ((False, "f8"), (True, "f8s")):
        v8, vsc, vm = OQ.per_channel_fp8(v, layout, 448.0, smooth)
        check_fp8(v8, vsc, vm, g, tag)


# ------------------------------------------------------------------------------------------------ GPU: kernels pinned
@pytest.fixture(scope="module")
def L():
    import lowbit_quant_fa2_paddle_b200 as L
    return L


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=NAMES)
def test_cuda_quantizers_match_reference_cuda_kernels(L, path):
    dev = torch.device("cuda:0")
    g, layout, dt, sm = load(path)
    q, k, v, km = (g[n].to(dev) for n in ("q", "k", "v", "km"))
    check_per_block(L.per_block_int8_cuda(q, k, km=km, sm_scale=sm, tensor_layout=layout), g, "pb", "pb")
    got = L.per_block_int8_cuda(q, k, km=None, sm_scale=sm, tensor_layout=layout)
    assert torch.equal(got[2].cpu(), g["pbn_k_int8"]) and torch.equal(got[3].cpu(), g["pbn_k_scale"])
    check_per_block(L.per_warp_int8(q, k, km=km, tensor_layout=layout), g, "pw", "pw")
    assert torch.equal(L.sub_mean_given(v, g["sm_vm"].to(dev), layout).cpu(), g["sm_v"]), "sub_mean differs from the reference kernel"
    for smooth, tag in ((False, "f8"), (True, "f8s")):
        v8, vsc, vm = L.per_channel_fp8(v, tensor_layout=layout, smooth_v=smooth)
        check_fp8(v8, vsc, vm, g, tag)
