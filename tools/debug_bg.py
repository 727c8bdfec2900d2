import sys, torch
from util_nrw import build_system, cuda_train_step, port, rel_err, synth
P = synth.make_params(0)
cfg = synth.PathConfig(n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4)
batch = synth.make_rays(48, cfg, seed=11)
mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
s = build_system(P, cfg, precision="bf16x6", backend=1 if mode == "simt" else 0, chunk_rows=2048)
r = s["renderer"]
dev = "cuda"
b = {k: v.to(dev) for k, v in batch.items()}
bg = torch.zeros([1, 3], device=dev)
for it in range(3):
    with torch.no_grad():
        res = r.render(b["rays"], b["ts"], b["label"], perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5)
    torch.cuda.synchronize()
    print(it, "nograd color absmax", float(res["color"].abs().max()), "bg tensor", bg.cpu().tolist(), "ptr", hex(bg.data_ptr()))
res = r.render(b["rays"], b["ts"], b["label"], perturb_overwrite=0, background_rgb=bg, cos_anneal_ratio=0.5)
torch.cuda.synchronize()
print("grad color absmax", float(res["color"].abs().max()), "bg tensor", bg.cpu().tolist())
res = r.render(b["rays"], b["ts"], b["label"], perturb_overwrite=0, background_rgb=None, cos_anneal_ratio=0.5)
torch.cuda.synchronize()
print("bg None color absmax", float(res["color"].abs().max()))
