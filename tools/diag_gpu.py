"""GPU bring-up diagnostics: prints error tables instead of asserting (run stage by stage under gpurun).

    python tests/diag_gpu.py gemm|sdf|sampler|render|all [--precision bf16x3] [--backend tc|simt]
"""
import argparse
import json
import sys
import time
import traceback

import numpy as np
import torch

from util_nrw import build_system, cuda_train_step, gemm_test, port, rel_err, synth


def stage_gemm(args):
    torch.manual_seed(0)
    dev = "cuda"
    rows = []
    shapes = [(300, 512, 512), (1000, 64, 512), (777, 128, 640), (640, 256, 192), (4096, 512, 64), (129, 512, 512)]
    for backend, bname in ((1, "simt"), (0, "tc")):
        for (M, N, K) in shapes:
            A = torch.randn(M, K, device=dev)
            B = torch.randn(N, K, device=dev) / np.sqrt(K)
            bias = torch.randn(N, device=dev)
            ref = (A.double() @ B.double().T + bias.double()).float()
            for P in (1, 2, 3):
                try:
                    D = gemm_test(backend, P, 0, 1, A, B, bias, 0)
                    rows.append((bname, "kk", M, N, K, P, rel_err(D.cpu(), ref.cpu())))
                except Exception as e:  # noqa
                    rows.append((bname, "kk", M, N, K, P, "ERR " + str(e)[:80]))
        # weight-gradient form: D[M,N] = sum_k A[k,M] B[k,N]
        for (Ks, M, N, ks) in [(5000, 512, 512, 4), (3000, 128, 640, 3), (2048, 256, 192, 1), (1000, 512, 64, 2)]:
            A = torch.randn(Ks, M, device=dev)
            B = torch.randn(Ks, N, device=dev) / np.sqrt(Ks)
            ref = (A.double().T @ B.double()).float()
            for P in (1, 2, 3):
                try:
                    D = gemm_test(backend, P, 1, ks, A, B, None, 0)
                    rows.append((bname, "mn", M, N, Ks, P, rel_err(D.cpu(), ref.cpu())))
                except Exception as e:  # noqa
                    rows.append((bname, "mn", M, N, Ks, P, "ERR " + str(e)[:80]))
    for r in rows:
        print("GEMM", *r)


def stage_sdf(args, P):
    cfg = synth.PathConfig()
    s = build_system(P, cfg, precision=args.precision, backend=args.backend_id, chunk_rows=4096)
    torch.manual_seed(1)
    x = (torch.rand(5000, 3) * 2 - 1) * 0.9
    ref_out = port.sdf_forward(P, x)
    ref_g = port.sdf_gradient(P, x, create_graph=False)
    dirs = torch.nn.functional.normalize(torch.randn(5000, 3), dim=-1)
    a = torch.randn(5000, 48)
    ref_rgb = port.color_forward(P, x, ref_g, dirs, ref_out[:, 1:], a).detach()
    sdf = s["renderer"].sdf(x.cuda().reshape(-1, 1, 3)).cpu()
    print("SDF  sdf_query rel", rel_err(sdf.reshape(-1), ref_out[:, 0].detach()))
    rgb, sdf2, nrm = s["renderer"].engine.neuconw_forward(x.cuda(), dirs.cuda(), a.cuda())
    print("SDF  neuconw sdf rel", rel_err(sdf2.cpu(), ref_out[:, 0].detach()), "normal rel", rel_err(nrm.cpu(), ref_g.detach()),
          "rgb rel", rel_err(rgb.cpu(), ref_rgb))
    p4 = torch.randn(3000, 4) * 0.5
    d3 = torch.nn.functional.normalize(torch.randn(3000, 3), dim=-1)
    a3 = torch.randn(3000, 48)
    rd, rr = port.nerf_forward(P, p4, d3, a3)
    dens, rgbn = s["nerf"](p4.cuda(), d3.cuda(), a3.cuda())
    print("NERF density rel", rel_err(dens.cpu(), rd.detach()), "rgb rel", rel_err(rgbn.cpu(), rr.detach()))


def stage_sampler(args, P):
    for name, cfg, R, pert in (("c1", synth.C1, 64, 0), ("c2p", synth.PathConfig(perturb=1.0, **synth.BRANDENBURG), 64, 1)):
        s = build_system(P, cfg, precision=args.precision, backend=args.backend_id, chunk_rows=4096)
        batch = synth.make_rays(R, cfg, seed=21)
        noise = synth.make_perturb_noise(R, cfg.n_outside, seed=5) if pert else None
        extras = {}
        with torch.no_grad():
            port.render(P, cfg, batch["rays"], batch["ts"], batch["label"], perturb_overwrite=-1 if pert else 0,
                        background_rgb=torch.zeros(1, 3), cos_anneal_ratio=0.5, noise=noise, extras=extras)
        r = s["renderer"]
        if noise is not None:
            r._noise_hook = lambda R_, n_, dev: (noise[0].to(dev), noise[1].to(dev))
        rays = batch["rays"].cuda()
        o = ((rays[:, 0:3] - r.origin.to("cuda").float()) / r.radius).float().contiguous()
        near, far = (rays[:, 6:7] / r.radius).float(), (rays[:, 7:8] / r.radius).float()
        with torch.no_grad():
            S, z, zo, sd, _, _ = r.sparse_sampler(o, rays[:, 3:6].contiguous(), near, far, cfg.perturb if pert else 0)
        print("SAMPLER", name, "S", S, "z max abs err", float((z.cpu() - extras["z_vals"]).abs().max()),
              "z_out err", float((zo.cpu() - extras["z_vals_outside"]).abs().max()),
              "sample_dist err", float((sd.cpu() - extras["sample_dist"]).abs().max()))


def stage_render(args, P):
    for name, cfg, R in (("small", synth.PathConfig(n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4), 48),
                         ("c1", synth.C1, 40)):
        batch = synth.make_rays(R, cfg, seed=11)
        t0 = time.time()
        res_p, loss_p, grads_p = port.train_step(P, cfg, batch, perturb_overwrite=0)
        t_cpu = time.time() - t0
        s = build_system(P, cfg, precision=args.precision, backend=args.backend_id, chunk_rows=args.chunk_rows)
        res_c, loss_c, grads_c = cuda_train_step(s, cfg, batch, perturb_overwrite=0)
        print("RENDER", name, "loss", float(loss_c), float(loss_p), "cpu_s", round(t_cpu, 2))
        for k in res_p:
            a, b = res_c[k].numpy(), res_p[k].detach().numpy()
            print("  out", k, tuple(a.shape), "rel", rel_err(a, b) if a.shape == b.shape else "SHAPE %s" % (b.shape,))
        worst = []
        for k in sorted(grads_p):
            worst.append((rel_err(grads_c[k].numpy(), grads_p[k].numpy()), k, float(grads_p[k].abs().max())))
        worst.sort(reverse=True)
        for w in worst[:12]:
            print("  grad", w)
        print("  grad median rel", float(np.median([w[0] for w in worst])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stage")
    ap.add_argument("--precision", default="bf16x6")
    ap.add_argument("--backend", default="tc")
    ap.add_argument("--chunk_rows", type=int, default=2048)
    args = ap.parse_args()
    args.backend_id = 1 if args.backend == "simt" else 0
    print("== diag", args.stage, args.precision, args.backend, torch.cuda.get_device_name(0))
    P = synth.make_params(seed=0)
    stages = dict(gemm=lambda: stage_gemm(args), sdf=lambda: stage_sdf(args, P), sampler=lambda: stage_sampler(args, P),
                  render=lambda: stage_render(args, P))
    todo = list(stages) if args.stage == "all" else [args.stage]
    for st in todo:
        try:
            stages[st]()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            print("STAGE FAILED", st)
    print("== done")


if __name__ == "__main__":
    main()
