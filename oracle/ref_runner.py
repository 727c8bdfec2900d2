"""Drive the UNMODIFIED reference (``/root/reference`` or its verbatim copy ``oracle/_ref``) for timing.

TEST INFRASTRUCTURE: imported only by ``bench.py``'s ``--impl reference`` / ``cpu_baseline`` /
``reference_torch_gpu`` legs and by tests.  One ``RefRunner`` owns the reference's own ``NeuconW`` / ``NeRF`` /
``nn.Embedding`` / ``NeuconWRenderer`` / ``NeuconWLoss`` objects (constructed by ``oracle.make_golden.build_reference``)
and performs what ``NeuconWSystem.training_step`` + Lightning do per step on the hot path
(lightning_modules/neuconw_system.py:159-176,337-360; train.py:61; utils/__init__.py:30):
zero_grad -> render -> loss -> backward -> clip_grad_norm_(0.99) -> Adam(eps=1e-7).step().
"""
import warnings

import torch

from . import ref_import, synth


def available():
    return ref_import.available()


class RefRunner:
    def __init__(self, cfg, P=None, device="cpu", lr=2e-4, optimizer=True):
        from .make_golden import build_reference

        self.cfg = cfg
        self.device = torch.device(device)
        P = P if P is not None else synth.make_params(seed=0)
        m = build_reference(cfg, P)
        for k in ("neuconw", "nerf", "emb"):
            m[k].to(self.device)
        self.m = m
        self.params = [p for k in ("emb", "neuconw", "nerf") for p in m[k].parameters()]
        self.optimizer = torch.optim.Adam(self.params, lr=lr, eps=1e-7, weight_decay=0) if optimizer else None

    def train_step(self, batch, perturb_overwrite=-1):
        m, cfg = self.m, self.cfg
        for p in self.params:
            p.grad = None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = m["renderer"].render(batch["rays"], batch["ts"], batch["label"], perturb_overwrite=perturb_overwrite,
                                       background_rgb=torch.zeros([1, 3], device=self.device),
                                       cos_anneal_ratio=cfg.cos_anneal_ratio)
            loss = sum(m["loss"](res, batch["rgbs"]).values())
            with torch.no_grad():
                mse = ((res["color"].detach() - batch["rgbs"]) ** 2).mean()
                self.last_metrics = {"loss": loss.detach(), "psnr": -10.0 * torch.log10(mse),
                                     "eikonal": res["gradient_error"].detach().mean(), "s_val": res["s_val"].detach().mean()}
            loss.backward()
        if self.optimizer is not None:
            torch.nn.utils.clip_grad_norm_(self.params, 0.99)
            self.optimizer.step()
        return loss.detach()
