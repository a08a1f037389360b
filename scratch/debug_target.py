import sys; sys.path.insert(0, '.')
import torch
from tests.test_gpu_parity import make_case, sub
from tests import kernel_model as km
from tests.util import nerr
import pfs_neural_net_b200.gnn as g
dev = torch.device('cuda:0')
for spec in (dict(seed=6, F=4, S=33, T=3), dict(seed=6, F=4, S=33, T=12), dict(seed=6, F=10, S=33, T=3)):
    case = make_case(**spec)
    F = case["F"]
    p = {k: v.double() if v.is_floating_point() else v for k, v in sub(case["state"], "t_model.").items()}
    ei = case["edge_index"]
    out64, saved = km.target_fwd(p, case["x_s"], case["x_t"], ei[0], ei[1], case["x_e"], case["u"], True, True, p["norm.running_mean"], p["norm.running_var"])
    mod = g.TModel(F); mod.load_state_dict(sub(case["state"], "t_model.")); mod = mod.to(dev).train()
    ins = [case[n].float().to(dev).requires_grad_(True) for n in ("x_s", "x_t", "x_e", "u")]
    out = mod(ins[0], ins[1], ei.to(dev), ins[2], ins[3])
    st = out.grad_fn.saved_tensors if out.grad_fn.name().startswith("TargetFunction") else out.grad_fn.next_functions[0][0].saved_tensors
    act_sum, y_pre, bn_save = st[-3], st[-2], st[-1]
    print(spec, "out", nerr(out, out64), "asum", nerr(act_sum[0], saved["asum"]), "y", nerr(y_pre[0], saved["y"]),
          "mu", nerr(bn_save[0, 0], saved["mu"]), "var", nerr(bn_save[0, 1], 1 / saved["r"] ** 2 - 1e-5))
    print("  y ours", y_pre[0].flatten()[:8].tolist()); print("  y ref ", saved["y"].flatten()[:8].tolist())
