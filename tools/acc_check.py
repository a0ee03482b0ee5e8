import os, sys, torch
sys.path.insert(0, "/root/repo")
import lowbit_quant_fa2_paddle_b200 as L
from oracle import attention as OA
torch.manual_seed(0)
dev = torch.device("cuda:0")
for (b,h,n,d,causal) in ((1,4,2048,64,False),(1,4,2048,128,True)):
    q,k,v = (torch.randn(b,h,n,d,dtype=torch.float16) for _ in range(3))
    k = k + 2*torch.randn(1,h,1,d).half()
    ref = OA.sdpa_fp32(q,k,v,"HND",causal)
    orc = OA.lowbit_fa_api(q,k,v,"HND",causal,compat_tail=False,pv_accum="fp32")
    o = L.lowbit_fa_qk_int8_pv_fp16_triton(q.to(dev),k.to(dev),v.to(dev),is_causal=causal).cpu()
    e1=(o.float()-orc.float()).abs(); e2=(o.float()-ref.float()).abs()
    print(f"variant {os.environ.get('LOWBIT_ATTN_VARIANT','0')} d={d} causal={causal}: vs oracle max {e1.max():.3e} mean {e1.mean():.3e} | vs sdpa32 max {e2.max():.3e} mean {e2.mean():.3e}")
