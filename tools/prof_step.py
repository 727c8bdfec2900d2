"""One training step inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200")); sys.path.insert(0, ROOT)
import torch
from nrw.synthetic import make_ray_batch
from nrw.train import TrainSystem
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--rays", type=int, default=8192)
ap.add_argument("--chunk_rows", type=int, default=32768)
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
s = TrainSystem(dev, precision=a.precision, chunk_rows=a.chunk_rows, batch_size=a.rays)
b = make_ray_batch(a.rays, seed=1, device=dev)
for _ in range(a.warm):
    s.training_step(b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
s.training_step(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
