"""The whole training step of reference src/train.py:136-141 (GNN forward, loss, backward, Adam) captured as a CUDA
graph: replays must reproduce the eager steps bit for bit, including the BatchNorm
running buffers and the sharpness schedule read from device memory."""
import copy

import pytest
import torch

from oracle import block_oracle as bo

pytestmark = pytest.mark.gpu


def _setup(dev, S=96, T=12, F=10):
    from pfs_neural_net_b200 import gnn as pg
    torch.manual_seed(0)
    model = pg.GNN(B=3, Fdim=F, T=T, F_s=1, F_t=2).to(dev).train()
    g = torch.Generator().manual_seed(1)
    class_info = torch.stack([0.5 + 3 * torch.rand(T, generator=g), 50 + 400 * torch.rand(T, generator=g)], 1).to(dev)
    graph = pg.BipartiteData(bo.complete_bipartite(S, T), torch.arange(S, dtype=torch.float32).reshape(-1, 1), class_info.cpu(),
                             2 + 8 * torch.rand(S * T, F, generator=g), torch.zeros(1, F))
    return model, graph, class_info


def test_graphed_train_step_matches_eager():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pfs_neural_net_b200.train_step import TrainStep
    dev = torch.device("cuda:0")
    model_a, graph, class_info = _setup(dev)
    model_b = copy.deepcopy(model_a)
    sharps = [0.5 + 0.1 * i for i in range(6)]
    runs = []
    for model, use_graph in ((model_a, False), (model_b, True)):
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
        init = copy.deepcopy(model.state_dict())
        step = TrainStep(model, graph, class_info, opt, use_graph=use_graph, warmup=2)
        # the warm-up / capture steps moved the weights: restart both runs from the same state
        model.load_state_dict(init)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True) if not use_graph else opt
        if use_graph:
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        else:
            step.opt = opt
        losses = []
        for i, sh in enumerate(sharps):
            torch.manual_seed(100 + i)                       # softfloor noise: same draw in both runs
            loss, util = step(sh)
            losses.append((float(loss), float(util)))
        runs.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
    (la, sa), (lb, sb) = runs
    for (a, ua), (b, ub) in zip(la, lb):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)) and abs(ua - ub) <= 1e-6 * max(1.0, abs(ua)), (la, lb)
    for k in sa:
        assert torch.allclose(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-6), k
    assert len({round(l[0], 3) for l in lb}) > 1             # the sharpness schedule reaches the kernels on replay
