"""Stage timing of a training step (CUDA events recorded in-stream by TrainSystem.stage_events)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200")); sys.path.insert(0, ROOT)
import torch
from nrw.synthetic import make_ray_batch
from nrw.train import TrainSystem
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3"); ap.add_argument("--rays", type=int, default=8192)
ap.add_argument("--chunk_rows", type=int, default=262144)
a = ap.parse_args()
dev = torch.device("cuda", 0)
s = TrainSystem(dev, precision=a.precision, chunk_rows=a.chunk_rows, batch_size=a.rays)
b = make_ray_batch(a.rays, seed=1, device=dev)
for _ in range(3): s.training_step(b)
torch.cuda.synchronize()
acc = {}
N = 5
t0 = time.perf_counter()
for _ in range(N):
    s.stage_events = []
    s.training_step(b)
    torch.cuda.synchronize()
    ev = s.stage_events
    for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
wall = (time.perf_counter() - t0) / N * 1e3
print("stage ms:", {k: round(v / N, 2) for k, v in acc.items()}, "sum", round(sum(acc.values()) / N, 2), "wall/step", round(wall, 2))
# host-side cost of issuing one step (no sync inside): time to return from training_step
s.stage_events = None
torch.cuda.synchronize(); t0 = time.perf_counter(); s.training_step(b); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host issue ms", round((t1 - t0) * 1e3, 2), "total ms", round((t2 - t0) * 1e3, 2))
