"""The whole training step of reference src/train.py:136-141 (GNN forward, loss, backward, Adam), row N2 of SURVEY.md 8f.

  * against the UNMODIFIED reference: tests/golden/train_steps.pt is the trajectory of /root/reference/src/train.py run as
    `__main__` for 16 epochs (the first 8 are compared, see tests/test_oracle_golden.py::_check_trajectory) (oracle/make_golden_train.py: small NFIBERS, seeded, softfloor noise recorded); TrainStep,
    eager and as a replayed CUDA graph, must reproduce its loss / utility curve and its final weights;
  * graph replay against eager execution of the repo's own step (bit-level agreement of the two launch modes);
  * constructing a TrainStep (warm-up + capture) must leave model, optimizer and generator untouched."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_steps.pt")


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _from_golden(dev):
    from pfs_neural_net_b200 import gnn as pg
    d = torch.load(GOLDEN)
    c = d["config"]
    model = pg.GNN(B=c["B"], Fdim=c["Fdim"], T=c["NCLASSES"], F_s=1, F_t=2)
    model.load_state_dict(d["init_state"], strict=True)
    model = model.to(dev).train()
    graph = pg.BipartiteData(d["edge_index"], d["x_s"], d["class_info"], d["x_e"], d["x_u"])
    return d, c, model, graph


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_reproduces_reference_trajectory(use_graph):
    from pfs_neural_net_b200.train_step import TrainStep
    dev = _dev()
    d, c, model, graph = _from_golden(dev)
    class_info = d["class_info"].to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=c["lr"], capturable=True)
    noise = torch.zeros(d["noise"].shape[1], device=dev)
    step = TrainStep(model, graph, class_info, opt, pclass=c["pclass"], pfiber=c["pfiber"], nfields=c["NFIELDS"],
                     total_time=c["TOTAL_TIME"], wutils=c["wutils"], wvar=c["wvar"], use_graph=use_graph, warmup=2, noise=noise)
    losses, utils = [], []
    for k in range(d["compare_epochs"]):
        noise.copy_(d["noise"][k])
        loss, util = step(float(d["sharps"][k]))
        losses.append(float(loss))
        utils.append(float(util))
    from tests.test_oracle_golden import _check_trajectory
    print("loss curve ours %s\n           ref  %s" % (["%.4f" % x for x in losses], ["%.4f" % x for x in d["losses"].tolist()]))
    _check_trajectory(d, losses, utils, {k: p.detach() for k, p in model.named_parameters()})
    per_epoch = {k: int(v) // c["nepochs"] for k, v in d["final_state"].items() if k.endswith("num_batches_tracked")}
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):          # 2 per epoch for the twice-applied edge norm, 1 otherwise
            assert int(v) == per_epoch[k] * d["compare_epochs"], k


def test_constructing_a_train_step_changes_nothing():
    from pfs_neural_net_b200.train_step import TrainStep
    dev = _dev()
    d, c, model, graph = _from_golden(dev)
    opt = torch.optim.Adam(model.parameters(), lr=c["lr"], capturable=True)
    before = copy.deepcopy(model.state_dict())
    torch.manual_seed(5)
    rng = torch.cuda.get_rng_state(dev).clone()
    TrainStep(model, graph, d["class_info"].to(dev), opt, use_graph=True, warmup=2)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert torch.equal(torch.cuda.get_rng_state(dev), rng)
    for st in opt.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                assert float(v.abs().max()) == 0.0, k          # moments and step counters back to "never stepped"


def test_graph_replay_matches_eager_steps():
    from pfs_neural_net_b200.train_step import TrainStep
    dev = _dev()
    runs = []
    for use_graph in (False, True):
        d, c, model, graph = _from_golden(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
        step = TrainStep(model, graph, d["class_info"].to(dev), opt, use_graph=use_graph, warmup=2)
        losses = []
        for i in range(6):
            torch.manual_seed(100 + i)                       # softfloor noise: same draw in both runs
            loss, util = step(0.5 + 0.1 * i)
            losses.append((float(loss), float(util)))
        runs.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
    (la, sa), (lb, sb) = runs
    for (a, ua), (b, ub) in zip(la, lb):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)) and abs(ua - ub) <= 1e-6 * max(1.0, abs(ua)), (la, lb)
    for k in sa:
        assert torch.allclose(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-6), k
    assert len({round(l[0], 3) for l in lb}) > 1             # the sharpness schedule reaches the kernels on replay
