import torch, sys
sys.path.insert(0,'/root/repo')
from probpose_code_b200 import _lib, ops
prec=_lib.PREC_FP16X3
for (m,n,k) in [(640,384,384),(640,384,1536),(640,384,3456),(1024,256,8192)]:
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) * 0.05
    ao, wo = ops.to_operand(a, prec), ops.to_operand(w, prec)
    ref = a.double() @ w.double().t()
    def dec(buf, rows, k):
        h = buf.view(torch.float16).view(rows, 2 * k).double()
        return (h[:, :k] + h[:, k:]) / 64.0
    refq = dec(ao,m,k) @ dec(wo,n,k).t()
    for tn in (128,192,256):
        out = ops.gemm(ao, wo, m, n, k, prec, tile_n=tn).double()
        print(m,n,k,tn,'rel vs exact %.3e'%((out-ref).abs().max()/ref.abs().max()).item(), 'vs quantized %.3e'%((out-refq).abs().max()/refq.abs().max()).item(), 'mean signed %.3e'%((out-refq).mean()/refq.abs().mean()).item())
