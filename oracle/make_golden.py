"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Run:  python -m oracle.make_golden
Imports /root/reference through ``oracle.ref_import`` (third-party imports stubbed),
loads ``oracle.synth.make_params`` weights into the reference's own ``NeuconW`` /
``NeRF`` / ``nn.Embedding`` modules, drives ``NeuconWRenderer.render`` +
``NeuconWLoss`` + ``backward`` on ``oracle.synth.make_rays`` batches and stores the
stage-boundary tensors.  Large parameter gradients are stored as seeded random
projections (``grad_probe``) + norms so the fixtures stay small.
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

from . import ref_import, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                          "golden")

SDF_CONFIG = dict(d_in=3, d_out=513, d_hidden=512, n_layers=8, skip_in=(4,), multires=6, bias=0.5,
                  scale=1, geometric_init=True, weight_norm=True, inside_outside=False)
COLOR_CONFIG = dict(d_in=9, d_feature=512, mode="idr", d_out=3, d_hidden=256, n_layers=4,
                    head_channels=128, static_head_layers=2, weight_norm=True, multires_view=4)


def install_injected_hits(renderer, hits, device="cpu"):
    """Config C3 without Kaolin: replace ONLY the Kaolin-backed ``get_near_far`` symbol that
    ``rendering/renderer.py`` imported (tools/prepare_data/generate_voxel.py:311) by a function returning injected
    trace results, and hand the renderer two placeholder octree dicts.  The reference's own
    ``get_near_far_octree`` / ``get_near_far_sdf`` / ``sparse_sampler`` arithmetic (renderer.py:380-568) runs unmodified."""
    import rendering.renderer as rr  # type: ignore

    coarse_tag, fine_tag = object(), object()

    def fake_get_near_far(rays_o, rays_d, octree, *a, **k):
        if octree is fine_tag:
            return hits["surface"].to(rays_o.device).clone(), None
        assert octree is coarse_tag
        return hits["sfm_near"].to(rays_o.device).clone(), hits["sfm_far"].to(rays_o.device).clone()

    rr.get_near_far = fake_get_near_far
    od = lambda tag, extra: dict(octree=tag, scene_origin=torch.zeros(3), scale=1.0, level=1, spc_data=None, **extra)
    renderer.octree_data = od(coarse_tag, {})
    renderer.fine_octree_data = od(fine_tag, {"voxel_size": hits["fine_voxel_sfm"]})
    renderer.nerf_far_override = True
    renderer.voxel_size = hits["voxel_size"]


def build_reference(cfg: synth.PathConfig, P):
    """Construct the reference modules and load the synthetic parameters into them."""
    ref = ref_import.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        neuconw = ref.NeuconW(sdfNet_config=SDF_CONFIG, colorNet_config=COLOR_CONFIG,
                              SNet_config=dict(init_val=0.3), in_channels_a=cfg.n_a, encode_a=True)
        nerf = ref.NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4,
                        skips=[4], encode_appearance=True, in_channels_a=cfg.n_a,
                        in_channels_dir=27, use_viewdirs=True)
    emb = torch.nn.Embedding(cfg.n_vocab, cfg.n_a)
    neuconw.load_state_dict({k[len("neuconw."):]: v for k, v in P.items() if k.startswith("neuconw.")})
    nerf.load_state_dict({k[len("nerf."):]: v for k, v in P.items() if k.startswith("nerf.")})
    emb.load_state_dict({"weight": P["embedding_a.weight"]})
    scene = tempfile.mkdtemp(prefix="nrw_scene_")
    import yaml
    with open(os.path.join(scene, "config.yaml"), "w") as f:
        yaml.safe_dump(dict(origin=[float(x) for x in cfg.origin], radius=float(cfg.radius),
                            sfm2gt=np.eye(4).tolist(),
                            eval_bbx=[[-cfg.radius] * 3, [cfg.radius] * 3]), f)
    renderer = ref.NeuconWRenderer(
        nerf=nerf, neuconw=neuconw, embeddings={"a": emb}, n_samples=cfg.n_samples,
        s_val_base=cfg.s_val_base, n_importance=cfg.n_importance, n_outside=cfg.n_outside,
        up_sample_steps=cfg.up_sample_steps, perturb=cfg.perturb, origin=list(cfg.origin),
        radius=cfg.radius, render_bg=cfg.render_bg, mesh_mask_list=cfg.mesh_mask_list,
        floor_normal=False, floor_labels=["road"], depth_loss=cfg.depth_loss,
        spc_options=dict(voxel_size=0.1, recontruct_path=scene, min_track_length=0),
        sample_range=cfg.sample_range, boundary_samples=cfg.boundary_samples,
        nerf_far_override=False, trim_sphere=cfg.trim_sphere)
    config = types.SimpleNamespace(NEUCONW=types.SimpleNamespace(
        MESH_MASK_LIST=cfg.mesh_mask_list, DEPTH_LOSS=cfg.depth_loss, FLOOR_NORMAL=False))
    loss = ref.NeuconWLoss(coef=1.0, igr_weight=cfg.igr_weight, mask_weight=cfg.mask_weight,
                           depth_weight=cfg.depth_weight, floor_weight=0.01, config=config)
    return dict(neuconw=neuconw, nerf=nerf, emb=emb, renderer=renderer, loss=loss)


def reference_train_step(cfg, P, batch, perturb_overwrite=0, rand_seed=None, hits=None):
    """The reference's own forward/loss/backward (NeuconWSystem.forward semantics,
    lightning_modules/neuconw_system.py:159-176,337-360)."""
    m = build_reference(cfg, P)
    if hits is not None:
        install_injected_hits(m["renderer"], hits)
    if rand_seed is not None:
        torch.manual_seed(rand_seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = m["renderer"].render(batch["rays"], batch["ts"], batch["label"],
                                   perturb_overwrite=perturb_overwrite,
                                   background_rgb=torch.zeros([1, 3]),
                                   cos_anneal_ratio=cfg.cos_anneal_ratio)
        loss_d = m["loss"](res, batch["rgbs"])
        loss = sum(loss_d.values())
        loss.backward()
    grads = {}
    for prefix, mod in (("neuconw.", m["neuconw"]), ("nerf.", m["nerf"]), ("embedding_a.", m["emb"])):
        for k, p in mod.named_parameters():
            grads[prefix + k] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
    return res, loss.detach(), grads, m


def grad_probe(grads, n_probe=4, seed=99):
    """Small digest of a gradient dict: per-parameter L2 norm + n_probe random projections."""
    out = {}
    for k in sorted(grads):
        g = grads[k].detach().double().reshape(-1)
        gen = torch.Generator().manual_seed(seed + (sum(map(ord, k)) % 100003))
        pr = torch.randn(n_probe, g.numel(), generator=gen, dtype=torch.float64)
        out[k] = torch.cat([g.norm().reshape(1), pr @ g]).numpy()
    return out


CASES = {
    # name: (cfg, n_rays, perturb_overwrite, torch seed for perturb draws)
    "small_det": (synth.PathConfig(n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4), 48, 0, None),
    "small_perturb": (synth.PathConfig(n_samples=16, n_importance=16, up_sample_steps=4, n_outside=4,
                                       perturb=1.0, **synth.BRANDENBURG), 48, -1, 123),
    "c1_slice": (synth.C1, 32, 0, None),
    # config C3: SfM-octree near/far override + surface-guided fine sampling + boundary samples with INJECTED octree
    # trace results (synth.make_injected_hits), perturbed strata, scene frame
    "fine_c3": (synth.PathConfig(n_samples=16, n_importance=16, up_sample_steps=4, n_outside=4, perturb=1.0,
                                 boundary_samples=10, sample_range=8.0, **synth.BRANDENBURG), 48, -1, 321),
}
FINE_CASES = {"fine_c3"}


def main():
    if not ref_import.available():
        print("reference tree not available; cannot regenerate golden vectors", file=sys.stderr)
        sys.exit(1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    P = synth.make_params(seed=0)
    for name, (cfg, n_rays, pov, rseed) in CASES.items():
        batch = synth.make_rays(n_rays, cfg, seed=11)
        hits = synth.make_injected_hits(batch, cfg) if name in FINE_CASES else None
        res, loss, grads, m = reference_train_step(cfg, P, batch, perturb_overwrite=pov, rand_seed=rseed, hits=hits)
        # re-run the sampler alone for z_vals (deterministic given the same torch seed)
        o = ((batch["rays"][:, 0:3] - torch.tensor(cfg.origin, dtype=torch.float64).float()) / cfg.radius).float()
        near, far = (batch["rays"][:, 6:7] / cfg.radius).float(), (batch["rays"][:, 7:8] / cfg.radius).float()
        if rseed is not None:
            torch.manual_seed(rseed)
        with torch.no_grad():
            _, z, z_out, sd = m["renderer"].sparse_sampler(
                o, batch["rays"][:, 3:6], near.clone(), far.clone(), cfg.perturb if pov < 0 else pov)
        arrays = {f"out.{k}": v.detach().numpy() for k, v in res.items()}
        arrays.update(z_vals=z.numpy(), z_vals_outside=z_out.numpy(), sample_dist=sd.numpy(),
                      loss=loss.numpy())
        for k, v in grad_probe(grads).items():
            arrays["gp." + k] = v
        for k in ("neuconw.deviation_network.variance", "neuconw.sdf_net.lin8.bias",
                  "neuconw.color_net.lin4.bias", "nerf.alpha_linear.weight", "nerf.rgb_linear.bias",
                  "neuconw.sdf_net.lin0.weight_g"):
            arrays["g." + k] = grads[k].numpy()
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"wrote {path}: loss={float(loss):.6f} {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
