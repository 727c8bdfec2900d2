"""Shared helpers for the GPU parity tests (CUDA path vs oracle port)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralrecon-w_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import neuconw_port as port  # noqa: E402
from oracle import synth  # noqa: E402

SDF_CONFIG = dict(d_in=3, d_out=513, d_hidden=512, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1,
                  geometric_init=True, weight_norm=True, inside_outside=False)
COLOR_CONFIG = dict(d_in=9, d_feature=512, mode="idr", d_out=3, d_hidden=256, n_layers=4, head_channels=128,
                    static_head_layers=2, weight_norm=True, multires_view=4)


def build_system(P, cfg, device="cuda", precision=None, backend=None, chunk_rows=None):
    """nrw modules + renderer carrying the synthetic parameters P (reference state_dict names)."""
    import nrw

    neuconw = nrw.NeuconW(SDF_CONFIG, COLOR_CONFIG, dict(init_val=0.3), in_channels_a=cfg.n_a, encode_a=True)
    nerf = nrw.NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4],
                    encode_appearance=True, in_channels_a=cfg.n_a, in_channels_dir=27, use_viewdirs=True)
    emb = torch.nn.Embedding(cfg.n_vocab, cfg.n_a)
    neuconw.load_state_dict({k[len("neuconw."):]: v for k, v in P.items() if k.startswith("neuconw.")})
    nerf.load_state_dict({k[len("nerf."):]: v for k, v in P.items() if k.startswith("nerf.")})
    emb.load_state_dict({"weight": P["embedding_a.weight"]})
    neuconw, nerf, emb = neuconw.to(device), nerf.to(device), emb.to(device)
    renderer = nrw.NeuconWRenderer(
        nerf=nerf, neuconw=neuconw, embeddings={"a": emb}, n_samples=cfg.n_samples, s_val_base=cfg.s_val_base,
        n_importance=cfg.n_importance, n_outside=cfg.n_outside, up_sample_steps=cfg.up_sample_steps,
        perturb=cfg.perturb, origin=list(cfg.origin), radius=cfg.radius, render_bg=cfg.render_bg,
        mesh_mask_list=cfg.mesh_mask_list, floor_normal=False, floor_labels=["road"], depth_loss=cfg.depth_loss,
        spc_options=dict(voxel_size=0.1, recontruct_path=None, min_track_length=0), sample_range=cfg.sample_range,
        boundary_samples=cfg.boundary_samples, nerf_far_override=False, trim_sphere=cfg.trim_sphere,
        precision=precision, gemm_backend=backend, chunk_rows=chunk_rows)
    return dict(neuconw=neuconw, nerf=nerf, emb=emb, renderer=renderer)


def cuda_train_step(sysd, cfg, batch, perturb_overwrite=0, noise=None):
    """NeuconWSystem.forward + NeuconWLoss + backward on the CUDA path. Returns (results, loss, grads)."""
    r = sysd["renderer"]
    for m in (sysd["neuconw"], sysd["nerf"], sysd["emb"]):
        m.zero_grad(set_to_none=True)
    if noise is not None:
        r._noise_hook = lambda R, n_out, dev: (noise[0].to(dev), noise[1].to(dev))
    dev = next(sysd["neuconw"].parameters()).device
    b = {k: v.to(dev) for k, v in batch.items()}
    res = r.render(b["rays"], b["ts"], b["label"], perturb_overwrite=perturb_overwrite,
                   background_rgb=torch.zeros([1, 3], device=dev), cos_anneal_ratio=cfg.cos_anneal_ratio)
    loss_d = port.loss_fn(cfg, res, b["rgbs"])
    loss = sum(loss_d.values())
    loss.backward()
    grads = {}
    for prefix, mod in (("neuconw.", sysd["neuconw"]), ("nerf.", sysd["nerf"]), ("embedding_a.", sysd["emb"])):
        for k, p in mod.named_parameters():
            grads[prefix + k] = (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)).cpu()
    return {k: v.detach().cpu() for k, v in res.items()}, loss.detach().cpu(), grads


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def gemm_test(backend, n_planes, mn_major, k_slices, A, B, bias=None, act=0):
    """D = A @ B^T (mn_major=0: A[M,K], B[N,K]; mn_major=1: A[K,M], B[K,N]) through the C ABI."""
    from nrw import _lib

    L = _lib.lib()
    if mn_major:
        K, M = A.shape
        N = B.shape[1]
    else:
        M, K = A.shape
        N = B.shape[0]
    D = torch.zeros(M, N, dtype=torch.float32, device=A.device)
    sb = L.nrw_gemm_test_scratch_bytes(M, N, K)
    scratch = torch.zeros(sb + 1024, dtype=torch.uint8, device=A.device)
    sp = (scratch.data_ptr() + 1023) // 1024 * 1024
    _lib.check(L.nrw_gemm_test(backend, n_planes, mn_major, k_slices, M, N, K, _lib.ptr(A.contiguous()),
                               _lib.ptr(B.contiguous()), _lib.ptr(bias), act, _lib.ptr(D), C.c_void_p(sp),
                               _lib.stream_ptr()), "nrw_gemm_test")
    torch.cuda.synchronize()
    return D
