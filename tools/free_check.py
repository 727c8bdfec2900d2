"""Does dropping a TrainSystem release its workspace? (bench.py builds other precision modes afterwards)"""
import gc, os, sys, weakref
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neuralrecon-w_b200"))
import torch
from nrw.train import TrainSystem
from nrw.synthetic import make_ray_batch
dev = torch.device("cuda:0")
gb = lambda: [round(x / 2**30, 1) for x in torch.cuda.mem_get_info(dev)]
print("start free/total", gb())
s = TrainSystem(dev, batch_size=8192)
b = make_ray_batch(8192, seed=1, device=dev)
for _ in range(2): l = s.training_step(b)
torch.cuda.synchronize()
print("after steps", gb(), "slots", s.renderer.engine.slots)
wr = weakref.ref(s.renderer.engine)
s = None
gc.collect(); torch.cuda.empty_cache()
print("after drop", gb(), "engine alive:", wr() is not None)
if wr() is not None:
    for r in gc.get_referrers(wr()):
        print("  referrer:", type(r), (list(r.keys())[:8] if isinstance(r, dict) else str(r)[:120]))
