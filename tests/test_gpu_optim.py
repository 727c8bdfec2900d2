"""Fused clip + Adam (nrw_grad_sumsq / nrw_adam_clip_step) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam
(the reference's optimiser path: train.py:61 gradient_clip_val=0.99, utils/__init__.py:30 Adam(eps=1e-7))."""
import numpy as np
import pytest
import torch

import util_nrw  # noqa: F401  (sys.path)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale", [1e-3, 5.0])      # below / above the clip threshold
def test_fused_clip_adam_matches_torch(scale):
    from nrw.train import FusedClipAdam
    torch.manual_seed(0)
    sizes = [1000003, 240000]
    ps = [torch.randn(n, device="cuda") for n in sizes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt_ref = torch.optim.Adam(ref, lr=2e-4, eps=1e-7, weight_decay=0)
    opt = FusedClipAdam(lr=2e-4, eps=1e-7, max_norm=0.99)
    for step in range(4):
        gs = [torch.randn(n, device="cuda") * scale * (1 + step) for n in sizes]
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 0.99)
        opt_ref.step()
        opt.step(list(zip(ps, gs)))
        for p, r in zip(ps, ref):
            # identical fp32 operation order; the only difference is the norm (double sum here, fp32 two-level in torch)
            assert (p - r.data).abs().max().item() <= 2e-7 * max(1.0, r.data.abs().max().item())
    assert opt.step_count == 4


def test_train_system_fused_equals_torch_and_variance_trains():
    from nrw.synthetic import make_ray_batch
    from nrw.train import TrainSystem
    dev = torch.device("cuda", 0)
    kw = dict(n_samples=16, n_importance=16, up_sample_steps=2, n_outside=4, batch_size=256, seed=3)
    a = TrainSystem(dev, fused_optimizer=True, **kw)
    b = TrainSystem(dev, fused_optimizer=False, **kw)
    batch = make_ray_batch(256, seed=5, device=dev)
    a.renderer.perturb = b.renderer.perturb = 0           # same (deterministic) strata on both systems
    v0 = float(a.neuconw.deviation_network.variance.detach())
    for _ in range(3):
        la = a.training_step(batch)
        lb = b.training_step(batch)
    assert abs(float(la) - float(lb)) <= 2e-4 * abs(float(lb))
    assert float(a.neuconw.deviation_network.variance) != v0            # gradient through the torch glue is kept
    pa = dict(a.neuconw.named_parameters()); pb = dict(b.neuconw.named_parameters())
    worst = 0.0
    for k in pb:
        d = (pa[k] - pb[k]).abs().max().item()
        worst = max(worst, d / (3 * 5e-6 + 1e-12))
    # three Adam steps of lr*(256/4096): per-parameter movement <= 3*lr; allow 2 % of it for sign flips of ~0 gradients
    lr = a.optimizer.param_groups[0]["lr"]
    for k in pb:
        assert (pa[k] - pb[k]).abs().max().item() <= 0.05 * 3 * lr + 1e-9, k
    assert (a.embedding_a.weight - b.embedding_a.weight).abs().max().item() <= 0.05 * 3 * lr + 1e-9
