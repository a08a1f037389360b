import sys, torch
sys.path.insert(0, ".")
from oracle import block_oracle as bo
from pfs_neural_net_b200 import gnn
import tests.test_gpu_wide_parity as tp
dev = torch.device("cuda:0")
kind, F, S, T, training, normed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), True, (len(sys.argv) < 6 or sys.argv[5] != "nonorm")
ei = tp._graph(kind, S, T, seed=F + S); E = ei.shape[1]
state = bo.random_block_state(F, seed=1)
state = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in state.items()}
if not normed: state = {k: v for k, v in state.items() if ".norm." not in k}
ins = tp._inputs(F, S, T, E, seed=7)
blk = gnn.Block(F, normed=normed).to(torch.bfloat16); blk.load_state_dict(state, strict=True); blk = blk.to(dev).train()
x = [t.to(dev).requires_grad_(True) for t in ins]
_, o_s, o_t, o_e, o_u = blk((ei.to(dev), *x))
gs = torch.Generator().manual_seed(11)
ups = [torch.randn(o.shape, generator=gs).bfloat16() for o in (o_s, o_t, o_e, o_u)]
torch.autograd.backward([o_s, o_t, o_e, o_u], [u.to(dev) for u in ups])
def run(dtype):
    sd = bo.cast_state(state, dtype)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    xx = [t.to(dtype).requires_grad_(True) for t in ins]
    outs = bo.block(sd, "", ei, *xx, training=True, normed=normed, buffers={}, rms_eps=tp.RMS_EPS_BF16)
    torch.autograd.backward(list(outs), [u.to(dtype) for u in ups])
    return sd, xx, outs
sd64, x64, o64 = run(torch.float64)
sd16, x16, o16 = run(torch.bfloat16)
def e(a, b): 
    a, b = a.detach().double().cpu(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
names = ("x_s", "x_t", "x_e", "u")
for n, a, r16, r in zip(names, (o_s, o_t, o_e, o_u), o16, o64): print("%-34s ours %.2e  ref-bf16 %.2e" % (n, e(a, r), e(r16, r)))
for n, a, r16, r in zip(names, x, x16, x64): print("%-34s ours %.2e  ref-bf16 %.2e" % ("g_" + n, e(a.grad, r.grad), e(r16.grad, r.grad)))
for k, p in blk.named_parameters():
    r = sd64[k].grad
    if r is None: continue
    print("%-34s ours %.2e  ref-bf16 %.2e   |ref| %.2e" % (k, e(p.grad, r), e(sd16[k].grad, r), r.abs().max().item()))
