import sys; sys.path.insert(0, '.')
import torch
from tests.test_gpu_parity import make_case, ours_module, oracle_module
from tests.util import nerr
dev = torch.device('cuda:0')
specs = [dict(seed=6, F=4, S=33, T=3), dict(seed=1), dict(seed=13, kind="sparse", S=45)]
if len(sys.argv) > 1:
    specs = specs[:int(sys.argv[1])]
for spec in specs:
    case = make_case(**spec)
    for name in ("edge_model", "s_model", "t_model", "global_model"):
        o, gin, gp, buf = ours_module(name, case, dev)
        o_ref, gin_ref, gp_ref, _ = oracle_module(name, case)
        torch.cuda.synchronize()
        print(spec, name, "fwd err %.2e" % nerr(o, o_ref), "gin", ["%.1e" % nerr(a, b) for a, b in zip(gin, gin_ref)])
