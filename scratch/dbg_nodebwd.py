import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_parity import make_case, oracle_module, ours_module
import tests.test_gpu_parity as tp
dev = torch.device('cuda:0')
for spec in [dict(seed=1), dict(seed=8, S=700), dict(seed=6, F=4, S=33, T=3), dict(seed=7, F=8, S=19, T=12)]:
    case = make_case(**spec)
    o_ref, gin_ref, gp_ref, _ = oracle_module('s_model', case)
    o, gin, gp, _ = ours_module('s_model', case, dev)
    F = case['F']
    for k in ('node_mlp_2.0.weight', 'node_mlp_2.2.weight', 'node_mlp_2.0.bias', 'node_mlp_2.2.bias'):
        a = gp[k].double().cpu(); b = gp_ref[k].double()
        print(spec, k, tuple(a.shape), 'err %.3e' % ((a - b).abs().max() / b.abs().max()).item(), 'mine max %.3e ref max %.3e' % (a.abs().max().item(), b.abs().max().item()))
    a = gp['node_mlp_2.0.weight'].double().cpu(); b = gp_ref['node_mlp_2.0.weight'].double()
    K9 = 9 * F
    a9, b9 = a[:, :K9], b[:, :K9]
    print('  first 9F cols err %.3e ; u cols err %.3e' % (((a9 - b9).abs().max() / b9.abs().max()).item(), ((a[:, K9:] - b[:, K9:]).abs().max() / b.abs().max()).item()))
    # per column-group-of-4 / row error pattern
    e = (a9 - b9).abs() / b9.abs().max()
    print('  rows with err>1e-3:', (e.max(1).values > 1e-3).nonzero().flatten().tolist()[:40])
    print('  cols with err>1e-3:', (e.max(0).values > 1e-3).nonzero().flatten().tolist()[:40])
    if a9.shape[0] == a9.shape[1]:
        pass
    print('  corr with ref: %.4f' % (torch.corrcoef(torch.stack([a9.flatten(), b9.flatten()]))[0, 1].item()))
