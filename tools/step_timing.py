"""Diagnose the timed loop of bench.py: async steps with / without the clock sampler, per-step sync, many steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200"))
import torch
import bench
from nrw.train import TrainSystem
from nrw.synthetic import make_ray_batch
w = dict(bench.WORKLOADS["C2"]); dev = torch.device("cuda:0")
sysm = TrainSystem(dev, n_samples=w["n_samples"], n_importance=w["n_importance"], up_sample_steps=w["up_sample_steps"],
                   n_outside=w["n_outside"], precision="bf16x3", chunk_rows=262144, batch_size=w["rays"], world_size=1, seed=66)
b = {k: v.to(dev) for k, v in make_ray_batch(w["rays"], seed=1).items()}
def timed(n, sync_each=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.time(); e0.record()
    for _ in range(n):
        l = sysm.training_step(b)
        if sync_each: l.item()
    t_host = time.time() - t0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_host / n * 1e3
for _ in range(3): sysm.training_step(b)
print("async 6 steps          : %.1f ms/step (host issue %.1f ms/step)" % timed(6))
print("sync each, 6 steps     : %.1f ms/step (host %.1f)" % timed(6, True))
cs = bench.ClockSampler(0); cs.start()
print("async 6 + nvml sampler : %.1f ms/step (host issue %.1f)" % timed(6))
print(cs.stop())
print("async 20 steps         : %.1f ms/step (host issue %.1f)" % timed(20))
print("mem", torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30)
# ---- pure host issue time of ONE step starting from an empty launch queue ----
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); l = sysm.training_step(b); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("one step: host issue %.1f ms, then GPU drain %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
import cProfile, pstats
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); sysm.training_step(b); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
