import sys, torch
from pfs_neural_net_b200 import wide_ops as wo
dev = torch.device("cuda:0")
E = 1 << 20
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
A = torch.randn(E, 128, device=dev).bfloat16(); B = torch.randn(512, 128, device=dev).bfloat16()
out = torch.empty(E, 512, device=dev, dtype=torch.bfloat16)
T = 512
t0 = torch.randn(E // T, 512, device=dev); t1 = torch.randn(T, 512, device=dev)
for _ in range(3):
    if mode == "plain":
        wo.gemm_nt(A, B, out_bf16=out, want="none")
    else:
        wo.gemm_nt(A, B, tab0=t0, div0=T, tab1=t1, mod1=T, act=True, out_bf16=out, want="none")
torch.cuda.synchronize()
