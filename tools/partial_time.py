import os, sys, torch
sys.path.insert(0, "/root/repo")
import lowbit_quant_fa2_paddle_b200 as L
from lowbit_quant_fa2_paddle_b200 import _native as NV, attention as A
dev = torch.device("cuda:0")
torch.manual_seed(0)
b,h,n,d = 1,32,8192,128
q,k,v = (torch.randn(b,h,n,d,dtype=torch.float16,device=dev) for _ in range(3))
km = L.k_mean(k)
qi,qs,ki,ks = L.per_block_q_int8_k_int4(q,k,km=km)
v8,vs,_ = L.per_channel_fp8(v, smooth_v=False)
def timed(f, reps=10):
    for _ in range(3): f()
    e0,e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
for causal in (False, True):
    st16 = A.forward_partial(None, qi, ki, v, qs, ks, "HND", causal, 0, 0, qk_mode=NV.QK_Q8K4)
    st8 = A.forward_partial(None, qi, ki, v8, qs, ks, "HND", causal, 0, 0, qk_mode=NV.QK_Q8K4, pv_mode=NV.PV_E4M3, v_scale=vs)
    t16 = timed(lambda: A.forward_partial(st16, qi, ki, v, qs, ks, "HND", causal, 0, 0, qk_mode=NV.QK_Q8K4))
    t8 = timed(lambda: A.forward_partial(st8, qi, ki, v8, qs, ks, "HND", causal, 0, 0, qk_mode=NV.QK_Q8K4, pv_mode=NV.PV_E4M3, v_scale=vs))
    f16 = timed(lambda: A._forward(qi, ki, v, qs, ks, "HND", torch.float16, False, causal, qk_mode=NV.QK_Q8K4))
    f8 = timed(lambda: A._forward(qi, ki, v8, qs, ks, "HND", torch.float16, False, causal, qk_mode=NV.QK_Q8K4, pv_mode=NV.PV_E4M3, v_scale=vs))
    print(f"causal={causal}: partial fp16 {t16:.3f} ms, partial fp8 {t8:.3f} ms | full fp16 {f16:.3f} ms, full fp8 {f8:.3f} ms")
