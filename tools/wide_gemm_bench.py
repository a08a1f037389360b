import torch, time
from pfs_neural_net_b200 import wide_ops as wo
dev = torch.device("cuda:0")
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
E = 1 << 20
for (M, N, K, name) in [(E, 512, 128, "E1"), (E, 128, 512, "E2"), (E, 256, 128, "S1"), (E, 256, 256, "S2"), (12500, 1280, 1152, "N3")]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: wo.gemm_nt(A, B, out_bf16=out, want="none"))
    fl = 2.0 * M * N * K; by = 2.0 * (M * K + M * N)
    print("NT %s M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (name, M, N, K, ms, fl / ms / 1e9, by / ms / 1e6))
    ms = bench(lambda: torch.matmul(A, B.T, out=out))
    print("   cuBLAS same: %.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9))
T = 512
t0 = torch.randn(E // T, 512, device=dev); t1 = torch.randn(T, 512, device=dev)
A = torch.randn(E, 128, device=dev).bfloat16(); B = torch.randn(512, 128, device=dev).bfloat16()
out = torch.empty(E, 512, device=dev, dtype=torch.bfloat16)
ms = bench(lambda: wo.gemm_nt(A, B, tab0=t0, div0=T, tab1=t1, mod1=T, act=True, out_bf16=out, want="none"))
print("NT E1 + tables + lrelu: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * E * 512 * 128 / ms / 1e9))
for (J, Kx, name) in [(512, 128, "dW1e"), (128, 512, "dW2"), (256, 256, "dW2s")]:
    D = torch.randn(E, J, device=dev).bfloat16(); X = torch.randn(E, Kx, device=dev).bfloat16()
    o = torch.empty(J, Kx, device=dev)
    ms = bench(lambda: wo.gemm_tn(D, X, out=o))
    fl = 2.0 * E * J * Kx; by = 2.0 * E * (J + Kx)
    print("TN %s E=%d J=%d K=%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (name, E, J, Kx, ms, fl / ms / 1e9, by / ms / 1e6))
    ms = bench(lambda: torch.matmul(D.T, X))
    print("   cuBLAS same: %.3f ms" % ms)
