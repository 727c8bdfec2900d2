"""Pins the CPU oracle port (oracle/neuconw_port.py) to golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py, run in the build container)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import neuconw_port as port
from oracle import synth
from oracle.make_golden import CASES, FINE_CASES, grad_probe

RTOL = 1e-4  # north-star tolerance (relative to the tensor's max magnitude)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_golden(name, params):
    cfg, n_rays, pov, rseed = CASES[name]
    G = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    batch = synth.make_rays(n_rays, cfg, seed=11)
    if rseed is not None:
        torch.manual_seed(rseed)
    hits = synth.make_injected_hits(batch, cfg) if name in FINE_CASES else None
    res, loss, grads = port.train_step(params, cfg, batch, perturb_overwrite=pov, hits=hits)
    assert abs(float(loss) - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    for k, v in res.items():
        g = G["out." + k]
        assert tuple(v.shape) == tuple(g.shape), k
        assert rel_err(v.detach().numpy(), g) < RTOL, k
    # integer-valued artefacts are exact
    assert np.array_equal(res["inside_sphere"].numpy(), G["out.inside_sphere"])
    gp = grad_probe(grads)
    for k, v in gp.items():
        assert rel_err(v, G["gp." + k]) < RTOL, k
    for k in G.files:
        if k.startswith("g."):
            assert rel_err(grads[k[2:]].numpy(), G[k]) < RTOL, k


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_sampler_matches_reference_golden(name, params):
    cfg, n_rays, pov, rseed = CASES[name]
    G = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    batch = synth.make_rays(n_rays, cfg, seed=11)
    if rseed is not None:
        torch.manual_seed(rseed)
    extras = {}
    with torch.no_grad():
        port.render(params, cfg, batch["rays"], batch["ts"], batch["label"], perturb_overwrite=pov,
                    background_rgb=torch.zeros(1, 3), cos_anneal_ratio=cfg.cos_anneal_ratio,
                    extras=extras, hits=synth.make_injected_hits(batch, cfg) if name in FINE_CASES else None)
    # same torch ops in the same order on the same machine class: bit-identical
    assert np.array_equal(extras["z_vals"].numpy(), G["z_vals"])
    assert np.array_equal(extras["z_vals_outside"].numpy(), G["z_vals_outside"])
    assert np.array_equal(extras["sample_dist"].numpy(), G["sample_dist"])


def test_result_dict_contract(params):
    """16 keys and shapes of rendering/renderer.py:899-916."""
    cfg = synth.PathConfig(n_samples=8, n_importance=8, up_sample_steps=2, n_outside=4)
    b = synth.make_rays(5, cfg, seed=3)
    with torch.no_grad():
        r = port.render(params, cfg, b["rays"], b["ts"], b["label"], perturb_overwrite=0,
                        background_rgb=torch.zeros(1, 3))
    S, T, R = 16, 20, 5
    want = dict(color=(R, 3), color_sphere=(R, 3), color_bg=(R, 3), s_val=(1, 1), cdf_fine=(R, S),
                gradients=(R, S, 3), mask_error=(R, 1), weights=(R, T), weights_sum=(R, 1),
                weights_max=(R, 1), gradient_error=(1,), inside_sphere=(R, S), depth=(R,),
                floor_normal_error=(R, 3), floor_y_error=(R, 3))
    assert set(r) == set(want) | {"sfm_depth_loss"}
    for k, s in want.items():
        assert tuple(r[k].shape) == s, k
