"""TEST INFRASTRUCTURE ONLY -- golden trajectory of the UNMODIFIED reference training loop.

Executes /root/reference/src/train.py as `__main__` (its source is compiled and run as is: reference
src/train.py:82-165 is the loop, :133-141 one step = zero_grad, forward, loss_function, backward, Adam) with
  * the shim packages of oracle/ref_shim first on sys.path (torch_scatter, torch_geometric, matplotlib),
  * a `config` module whose constants are overridden for a small, CPU, few-epoch run (NFIBERS, nepochs, device, paths),
  * torch seeded, and `torch.rand_like` (softfloor's noise, src/train.py:22) replaced by a recorded seeded draw so that
    the B200 path can be fed the same noise,
and commits the inputs, the initial weights, the per-epoch noise / sharpness / loss / utility and the final weights as
tests/golden/train_steps.pt.  tests/test_gpu_train_step.py replays it through pfs-neural-net_b200's TrainStep.

    python oracle/make_golden_train.py        (build container only: needs /root/reference)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

NFIBERS, NEPOCHS, SEED = 96, 16, 20250
COMPARE = 8          # epochs the parity tests compare (see tests/test_oracle_golden.py::_check_trajectory)


def main():
    src_dir = os.path.join(ref_loader.REFERENCE_ROOT, "src")
    tmp = tempfile.mkdtemp(prefix="pfs_golden_train_")
    os.makedirs(os.path.join(tmp, "params"))
    os.makedirs(os.path.join(tmp, "figures"))
    os.makedirs(os.path.join(tmp, "src"))
    with ref_loader._patched_path():
        import config as ref_config                      # the reference's constants ...
        cfg = types.ModuleType("config")
        cfg.__dict__.update({k: v for k, v in ref_config.__dict__.items() if not k.startswith("__")})
        cfg.device = torch.device("cpu")                 # ... overridden for a small CPU run
        cfg.NFIBERS, cfg.nepochs = NFIBERS, NEPOCHS
        cfg.datafile = os.path.join(ref_loader.REFERENCE_ROOT, "params", "increasing.txt")
        cfg.checkpoint_path = os.path.join(tmp, "params", "model_gnn_")
        sys.modules["config"] = cfg
        sys.modules.pop("gnn", None)
        os.chdir(os.path.join(tmp, "src"))               # '../figures/...' of the report lands in the temp dir
        noise_log = []
        gen = torch.Generator().manual_seed(SEED + 1)
        real_rand_like = torch.rand_like

        def recorded_rand_like(x, *a, **k):
            n = torch.rand(x.shape, generator=gen, dtype=x.dtype)
            noise_log.append(n.clone())
            return n

        torch.rand_like = recorded_rand_like
        # weights after COMPARE optimizer steps (a hook on Adam.step in this harness; train.py itself is untouched)
        snap = {}
        real_step = torch.optim.Adam.step
        calls = [0]

        def counting_step(self, *a, **k):
            r = real_step(self, *a, **k)
            calls[0] += 1
            if calls[0] == COMPARE:
                for group in self.param_groups:
                    snap["params"] = [p.detach().clone() for p in group["params"]]
            return r

        torch.optim.Adam.step = counting_step
        torch.manual_seed(SEED)
        g = {"__name__": "__main__", "__file__": os.path.join(src_dir, "train.py")}
        sys.argv = ["train.py"]
        with open(os.path.join(src_dir, "train.py")) as f:
            code = compile(f.read(), os.path.join(src_dir, "train.py"), "exec")
        try:
            exec(code, g)                                # the loop runs; the plots after it need matplotlib
        except Exception as e:                           # reporting / plotting past the final checkpoint
            print("reference train.py stopped after the training loop: %s: %s" % (type(e).__name__, str(e)[:100]))
        finally:
            torch.rand_like = real_rand_like
            torch.optim.Adam.step = real_step
        assert len(noise_log) == NEPOCHS, len(noise_log)
        losses, utils = np.array(g["losses"]), np.array(g["objective"])
        final = {k: v.detach().clone() for k, v in g["gnn"].state_dict().items()}
        mid = {n: t for (n, _), t in zip(g["gnn"].named_parameters(), snap["params"])}
        graph = g["graph"]
        # the initial weights: same seed, same order of draws as src/train.py:97-107 (x_e first, then GNN(...))
        import gnn as ref_gnn
        torch.manual_seed(SEED)
        x_e0 = 2.0 + (10.0 - 2.0) * torch.rand(size=(NFIBERS * cfg.NCLASSES, cfg.Fdim))
        assert torch.equal(x_e0, graph.x_e)
        model0 = ref_gnn.GNN(Fdim=cfg.Fdim, B=3, F_s=1, F_t=2, T=cfg.NCLASSES)
        init = {k: v.detach().clone() for k, v in model0.state_dict().items()}
    sharps = [cfg.sharps[0] + (cfg.sharps[1] - cfg.sharps[0]) * e / NEPOCHS for e in range(NEPOCHS)]
    out = {
        "config": dict(NFIBERS=NFIBERS, NCLASSES=cfg.NCLASSES, Fdim=cfg.Fdim, B=3, nepochs=NEPOCHS, lr=cfg.lr, pclass=cfg.pclass,
                       pfiber=cfg.pfiber, wutils=cfg.wutils, wvar=cfg.wvar, NFIELDS=cfg.NFIELDS, TOTAL_TIME=cfg.TOTAL_TIME,
                       seed=SEED),
        "class_info": graph.x_t.detach().clone(), "x_s": graph.x_s.detach().clone(), "x_e": graph.x_e.detach().clone(),
        "x_u": graph.x_u.detach().clone(), "edge_index": graph.edge_index.detach().clone(),
        "init_state": init, "final_state": final, "compare_epochs": COMPARE, "params_after_compare": mid, "noise": torch.stack(noise_log), "sharps": torch.tensor(sharps),
        "losses": torch.tensor(losses), "utilities": torch.tensor(utils),
    }
    path = os.path.join(ROOT, "tests", "golden", "train_steps.pt")
    torch.save(out, path)
    print("wrote %s: %d epochs, loss %.4f -> %.4f, utility %.5f -> %.5f" % (path, NEPOCHS, losses[0], losses[-1], utils[0], utils[-1]))


if __name__ == "__main__":
    main()
