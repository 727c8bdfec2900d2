"""GPU: engine state handling: workspace bounds vs point queries, generation-stamped forward cache,
RAY_MASK_LIST filter, weight packing once per parameter version, and the world_size=2 branch of
TrainSystem.training_step (two processes on ONE GPU, gloo collectives on CUDA tensors)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from util_nrw import build_system, rel_err, synth

pytestmark = pytest.mark.gpu


def _small_system(**kw):
    from nrw.train import TrainSystem

    return TrainSystem(torch.device("cuda", 0), n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4, n_vocab=64,
                       precision="bf16x3", chunk_rows=4096, batch_size=256, **kw)


def _batch(R, seed, device="cuda", labels=(0.0, 1.0, 2.0, 6.0)):
    from nrw.synthetic import make_ray_batch

    b = make_ray_batch(R, seed=seed, n_vocab=64, device=device)
    g = torch.Generator().manual_seed(seed)
    b["label"] = torch.tensor(labels)[torch.randint(0, len(labels), (R,), generator=g)].to(device)
    return b


def test_point_query_does_not_inflate_the_ray_workspace():
    sysm = _small_system()
    b = _batch(256, 3)
    sysm.training_step(b)
    eng = sysm.renderer.engine
    slots, bound, ws_bytes = eng.slots, eng.bound, eng.workspace.numel()
    assert slots[0] >= 1
    pts = torch.rand(1 << 20, 1, 3, device="cuda") * 2 - 1
    sdf = sysm.renderer.sdf(pts)                    # 1M-point query (octree refresh / mesh extraction size)
    assert sdf.shape == (1 << 20, 1) and torch.isfinite(sdf).all()
    assert eng.bound == bound and eng.slots == slots and eng.workspace.numel() == ws_bytes
    loss = sysm.training_step(b)
    assert torch.isfinite(loss) and eng.slots == slots


def test_backward_of_an_older_render_recomputes_instead_of_reusing_newer_activations():
    P = synth.make_params(seed=0)
    cfg = synth.PathConfig(n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4)
    s = build_system(P, cfg, precision="bf16x3", backend=0, chunk_rows=2048)
    r = s["renderer"]
    bA = {k: v.cuda() for k, v in synth.make_rays(40, cfg, seed=1).items()}
    bB = {k: v.cuda() for k, v in synth.make_rays(40, cfg, seed=2).items()}
    bg = torch.zeros(1, 3, device="cuda")

    def grads_of(render_first, also_second):
        for m in (s["neuconw"], s["nerf"], s["emb"]):
            m.zero_grad(set_to_none=True)
        resA = r.render(render_first["rays"], render_first["ts"], render_first["label"], perturb_overwrite=0,
                        background_rgb=bg, cos_anneal_ratio=0.5)
        if also_second is not None:     # a second grad-enabled render of the SAME shape overwrites the shared slots
            r.render(also_second["rays"], also_second["ts"], also_second["label"], perturb_overwrite=0,
                     background_rgb=bg, cos_anneal_ratio=0.5)
        (resA["color"].sum() + resA["gradient_error"].sum()).backward()
        return r.engine.last_flat_grad.clone()

    g_ref = grads_of(bA, None)
    g_two = grads_of(bA, bB)
    assert rel_err(g_two.cpu().numpy(), g_ref.cpu().numpy()) < 2e-4     # fp32 atomics reorder only


def test_ray_mask_list_filter_matches_reference():
    """neuconw_system.py:345-355: person / car / bicycle / minibike rays are removed before rendering."""
    from nrw.renderer import LABEL_IDS

    sysm = _small_system()
    labels = (0.0, 2.0, float(LABEL_IDS["person"]), float(LABEL_IDS["car"]), float(LABEL_IDS["bicycle"]), float(LABEL_IDS["minibike"]))
    b = _batch(300, 5, labels=labels)
    fb = sysm.filter_rays(b)
    keep = torch.ones(300, dtype=torch.bool, device="cuda")
    for name in ("person", "car", "bicycle", "minibike"):
        keep[LABEL_IDS[name] == b["label"]] = False
    assert 0 < int(keep.sum()) < 300
    for k in ("rays", "rgbs", "ts", "label"):
        assert torch.equal(fb[k], b[k][keep]), k
    loss = sysm.training_step(b)                  # variable R through the engine
    assert torch.isfinite(loss)


def test_weights_are_packed_once_per_parameter_version():
    from nrw import _lib

    sysm = _small_system()
    b = _batch(128, 7)
    sysm.training_step(b)
    eng = sysm.renderer.engine
    # pack() must be a no-op while the parameter version token is unchanged
    with torch.no_grad():
        sysm.forward(b["rays"], b["ts"], b["label"])         # first use after the optimizer step: packs
    tok0 = eng.packed_version
    n0 = eng.L.nrw_launch_count()
    with torch.no_grad():
        sysm.forward(b["rays"], b["ts"], b["label"])
        n_fwd = eng.L.nrw_launch_count() - n0
        sysm.renderer.sdf(torch.zeros(10, 1, 3, device="cuda"))
        n1 = eng.L.nrw_launch_count()
        sysm.forward(b["rays"], b["ts"], b["label"])
        assert eng.L.nrw_launch_count() - n1 == n_fwd        # identical launch count: no hidden repack
    assert eng.packed_version == tok0
    sysm.training_step(b)                      # optimizer step -> new version -> repack on next use
    with torch.no_grad():
        sysm.forward(b["rays"], b["ts"], b["label"])
    assert eng.packed_version != tok0
    # a torch in-place update of a parameter view also invalidates the packed copy
    tok1 = eng.packed_version
    with torch.no_grad():
        sysm.neuconw.sdf_net.lin1.bias.add_(0.01)
        before = sysm.renderer.sdf(torch.zeros(4, 1, 3, device="cuda") + 0.3).clone()
    assert eng.packed_version != tok1
    with torch.no_grad():
        sysm.neuconw.sdf_net.lin1.bias.sub_(0.01)
        after = sysm.renderer.sdf(torch.zeros(4, 1, 3, device="cuda") + 0.3)
    assert not torch.equal(before, after)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join({root!r}, "neuralrecon-w_b200")); sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_gpu_engine_state import _small_system, _batch
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
sysm = _small_system(world_size=2)
sysm.renderer.perturb = 0.0      # deterministic strata: the single-process restatement must see the same samples
for step in range(2):
    sysm.training_step(_batch(96, 100 + 10 * step + rank))
eng = sysm.renderer.engine
torch.save(dict(flat=eng.flat.cpu(), emb=sysm.embedding_a.weight.data.cpu()), {out!r} + f".{{rank}}")
dist.barrier()
dist.destroy_process_group()
"""


def test_train_system_two_ranks_equals_gradient_mean(tmp_path):
    """world_size=2 branch of TrainSystem.training_step: two processes (both on cuda:0, gloo all-reduce of the CUDA
    gradient buffers) vs ONE process that averages the two per-rank gradients itself and applies the same fused
    clip+Adam update with the same world-scaled learning rate."""
    out = str(tmp_path / "params")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    logs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)[-3000:]
    got = [torch.load(out + f".{r}") for r in range(2)]
    assert torch.equal(got[0]["flat"], got[1]["flat"]) and torch.equal(got[0]["emb"], got[1]["emb"])   # replicas stay identical
    # single-process restatement of DDP: mean of the per-rank gradients, one optimizer step per global step
    sysm = _small_system(world_size=2)
    sysm.world_size = 1                        # no process group here; the learning rate was scaled for world 2
    sysm.renderer.perturb = 0.0
    for step in range(2):
        _, f0, e0 = sysm.compute_grads(_batch(96, 100 + 10 * step + 0))
        f0, e0 = f0.clone(), e0.clone()
        _, f1, e1 = sysm.compute_grads(_batch(96, 100 + 10 * step + 1))
        f1.add_(f0).div_(2)
        e1.add_(e0).div_(2)
        sysm.apply_grads(f1, e1)
    ref_flat, ref_emb = sysm.renderer.engine.flat.cpu(), sysm.embedding_a.weight.data.cpu()
    lr = sysm.optimizer.param_groups[0]["lr"]
    # Adam's first steps move a weight by ~lr * sign(g): an element whose gradient sits at the fp32 noise floor can take
    # the other sign under a different atomic-add order, so compare robustly: almost every element within 5 % of lr
    for name, a, b_ in (("flat", got[0]["flat"], ref_flat), ("emb", got[0]["emb"], ref_emb)):
        diff = (a - b_).abs()
        frac_bad = float((diff > 0.05 * lr).float().mean())
        assert frac_bad < 1e-3, (name, frac_bad, float(diff.max()), lr)
        assert float(diff.mean()) < 0.01 * lr, (name, float(diff.mean()), lr)
