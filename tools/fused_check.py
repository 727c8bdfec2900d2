"""Fused forward-only SDF chain (NRW_SDF_FUSED) vs the per-layer chain: writes sdf of seeded points to a file (run once per setting,
the switch is read once per process), or compares two such files."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200")); sys.path.insert(0, ROOT)
import torch

if sys.argv[1] == "cmp":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    d = (a - b).abs()
    print("n", a.numel(), "max|sdf|", float(a.abs().max()), "max abs diff", float(d.max()), "mean abs diff", float(d.mean()),
          "n_diff", int((d > 0).sum()), "finite", bool(torch.isfinite(b).all()))
    bad = int((d > 2e-6 * float(a.abs().max()) + 1e-7).sum())
    print("rows beyond 2e-6 rel:", bad)
    if bad:
        idx = torch.nonzero(d > 2e-6 * float(a.abs().max()) + 1e-7).reshape(-1)[:10]
        print("first bad rows", idx.tolist(), a[idx].tolist(), b[idx].tolist())
    sys.exit(1 if bad else 0)

from nrw.train import TrainSystem
n = int(sys.argv[2])
dev = torch.device("cuda", 0)
s = TrainSystem(dev, precision="mixed", chunk_rows=262144)
g = torch.Generator().manual_seed(5)
pts = (torch.rand(n, 1, 3, generator=g) * 2 - 1).to(dev)
with torch.no_grad():
    out = s.renderer.sdf(pts).reshape(-1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        s.renderer.sdf(pts)
    e1.record(); torch.cuda.synchronize()
print("fused" if os.environ.get("NRW_SDF_FUSED", "1") != "0" else "per-layer", "n", n, "ms per query", e0.elapsed_time(e1) / 3, "Mq/s", n / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e6)
torch.save(out.cpu(), sys.argv[1])
