"""Live comparison of the oracle port with the unmodified reference (build container only;
skipped where /root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import neuconw_port as port
from oracle import ref_import, synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_port_vs_reference_live(params):
    from oracle.make_golden import reference_train_step

    cfg = synth.PathConfig(n_samples=12, n_importance=12, up_sample_steps=3, n_outside=6,
                           s_val_base=2, cos_anneal_ratio=0.25, **synth.BRANDENBURG)
    batch = synth.make_rays(24, cfg, seed=5)
    res_r, loss_r, grads_r, _ = reference_train_step(cfg, params, batch, perturb_overwrite=0)
    res_p, loss_p, grads_p = port.train_step(params, cfg, batch, perturb_overwrite=0)
    assert abs(float(loss_r) - float(loss_p)) < 1e-5 * abs(float(loss_r))
    for k in res_r:
        a, b = res_p[k].detach().numpy(), res_r[k].detach().numpy()
        assert a.shape == b.shape, k
        if a.size:
            assert np.abs(a - b).max() <= 1e-4 * (np.abs(b).max() + 1e-12), k
    for k in grads_r:
        a, b = grads_p[k].numpy(), grads_r[k].numpy()
        assert np.abs(a - b).max() <= 1e-4 * (np.abs(b).max() + 1e-12), k


def test_state_dict_names_match_reference(params):
    """oracle.synth parameter names/shapes == reference checkpoint layout."""
    from oracle.make_golden import build_reference

    m = build_reference(synth.PathConfig(), params)
    names = {}
    for pre, mod in (("neuconw.", m["neuconw"]), ("nerf.", m["nerf"]), ("embedding_a.", m["emb"])):
        for k, v in mod.state_dict().items():
            names[pre + k] = tuple(v.shape)
    assert names == {k: tuple(v.shape) for k, v in params.items()}
