"""Device-resident ray cache: the reference's training data path for the ray-cache mode
(datasets/data.py:83-119 per-rank split assignment, datasets/phototourism.py:467-515 cache loading, :709-724
__getitem__, torch DataLoader(shuffle=True, batch_size) batching) with the RAY_MASK_LIST filter of
lightning_modules/neuconw_system.py:345-355 fused into the gather.

The reference keeps the ~40 GB cache as a CPU tensor and lets a Python DataLoader index it item by item; at
>= 50 k rays/s per GPU that cannot keep up.  Here the rank's shard lives in HBM in the reference's on-disk layout
(rays [n,12] float32 = o3, d3, near, far, ts, label, depth, weight; rgbs [n,3]), a device-side permutation plays
RandomSampler, and ONE call of nrw_raycache_gather (csrc/dataio.cu) produces the batch dict of training_step -
already filtered and compacted.  The batch for step i+1 is prepared on a side stream while step i runs, and its
row count reaches the host through a pinned buffer, so the training stream never waits for the loader."""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import NrwError, check, ptr

RAY_MASK_LIST = ("person", "car", "bicycle", "minibike")      # config/train_brandenburg_gate.yaml:27


def local_split(items, world_size, rank, seed=6):
    """Split names this rank loads (DataModule._get_local_split, datasets/data.py:83-100): seeded permutation, topped up to a
    multiple of the world size by a seeded draw with replacement, equal contiguous runs per rank."""
    names = list(items)
    order = np.random.RandomState(seed).permutation(names)
    shortfall = (-len(names)) % world_size
    if shortfall:
        order = np.concatenate([order, np.random.RandomState(seed).choice(names, shortfall, replace=True)])
    run = len(order) // world_size
    return list(order[run * rank: run * (rank + 1)])


def load_split_arrays(root_dir, split_path, split_names, img_downscale=1):
    """PhototourismDataset cache loading (datasets/phototourism.py:467-515): concatenates
    <root>/<split_path>/<split>/rays{d}.npz|h5 and rgbs{d}.npz|h5 of the assigned splits (host numpy arrays)."""
    rays, rgbs = [], []
    first = os.path.join(root_dir, split_path, split_names[0])
    cache_type = sorted(os.listdir(first))[0].split(".")[-1]
    for name in split_names:
        for arr, key, out in (("rays", "rays", rays), ("rgbs", "rgbs", rgbs)):
            path = os.path.join(root_dir, split_path, name, f"{arr}{img_downscale}.{cache_type}")
            if cache_type == "npz":
                out.append(np.load(path)["arr_0"])
            elif cache_type == "h5":
                try:
                    import h5py
                except ImportError as e:      # noqa
                    raise NrwError("ray cache is stored as .h5 and h5py is not installed; re-export it as .npz "
                                   "(tools/prepare_data/prepare_data_cache.py --cache_type npz)") from e
                with h5py.File(path, "r") as f:
                    out.append(f[key][:])
            else:
                raise NrwError(f"unknown ray-cache file type {cache_type!r} in {first}")
    return np.concatenate(rays, 0), np.concatenate(rgbs, 0)


class RayCache:
    """rays [n,12], rgbs [n,3] (torch / numpy, any device) -> resident shard on `device`.

    next_batch() returns {"rays" [m,10], "rgbs" [m,3], "ts" [m] int64, "label" [m], "n_valid": m} where m <= batch_size
    rows survived the label filter; epochs follow RandomSampler semantics (a fresh permutation per epoch, the last
    batch of an epoch is short)."""

    def __init__(self, rays, rgbs, batch_size, device, ray_mask_list=RAY_MASK_LIST, seed=0, prefetch=True, drop_last=False):
        from .renderer import LABEL_IDS

        self.device = torch.device(device)
        self.rays = torch.as_tensor(rays, dtype=torch.float32).to(self.device).contiguous()
        self.rgbs = torch.as_tensor(rgbs, dtype=torch.float32).to(self.device).contiguous()
        if self.rays.dim() != 2 or self.rays.shape[1] != 12 or self.rgbs.shape != (self.rays.shape[0], 3):
            raise NrwError(f"RayCache: expected rays [n,12] and rgbs [n,3] (cache with semantics), got "
                           f"{tuple(self.rays.shape)} / {tuple(self.rgbs.shape)}")
        if not self.rays.is_cuda:
            raise NrwError("RayCache: the shard must live on a CUDA device (no CPU path)")
        self.n = self.rays.shape[0]
        self.batch_size = int(batch_size)
        self.drop_last = bool(drop_last)
        ids = [LABEL_IDS[name] for name in (ray_mask_list or ())]
        self.mask = (C.c_int32 * max(len(ids), 1))(*ids) if ids else (C.c_int32 * 1)(0)
        self.n_mask = len(ids)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        self.L = _lib.lib()
        self.stream = torch.cuda.Stream(device=self.device) if prefetch else None
        B = self.batch_size
        sb = self.L.nrw_compact_scratch_bytes(B)
        self._bufs = []
        for _ in range(2):
            with torch.cuda.device(self.device):
                self._bufs.append(dict(
                    rays=torch.empty(B, 10, dtype=torch.float32, device=self.device),
                    rgbs=torch.empty(B, 3, dtype=torch.float32, device=self.device),
                    ts=torch.empty(B, dtype=torch.int64, device=self.device),
                    label=torch.empty(B, dtype=torch.float32, device=self.device),
                    n_valid=torch.zeros(1, dtype=torch.int64, device=self.device),
                    n_host=torch.zeros(1, dtype=torch.int64).pin_memory(),
                    scratch=torch.empty(sb + 256, dtype=torch.uint8, device=self.device),
                    event=torch.cuda.Event()))
        self.perm = None
        self.cursor = 0
        self.epoch = 0
        self._slot = 0
        self._pending = None
        if prefetch:
            self._pending = self._enqueue()

    def __len__(self):                       # batches per epoch, as len(DataLoader)
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    # ------------------------------------------------------------------------------------------------------------
    def _next_indices(self):
        if self.perm is None or self.cursor >= self.n or (self.drop_last and self.cursor + self.batch_size > self.n):
            self.perm = torch.randperm(self.n, generator=self.gen, device=self.device)
            self.cursor = 0
            self.epoch += 1
        idx = self.perm[self.cursor:self.cursor + self.batch_size]
        self.cursor += idx.shape[0]
        return idx

    def gather(self, index, buf=None):
        """Batch for an explicit index vector (int64, device): __getitem__ over the vector + label filter."""
        buf = buf or self._bufs[0]
        B = int(index.shape[0])
        if B > self.batch_size:
            raise NrwError(f"RayCache.gather: {B} indices exceed the batch size {self.batch_size}")
        sp = (buf["scratch"].data_ptr() + 255) // 256 * 256
        check(self.L.nrw_raycache_gather(ptr(self.rays), ptr(self.rgbs), self.n, ptr(index), B, self.mask, self.n_mask,
                                         ptr(buf["rays"]), ptr(buf["rgbs"]), ptr(buf["ts"]), ptr(buf["label"]),
                                         ptr(buf["n_valid"]), C.c_void_p(sp), _lib.stream_ptr()), "nrw_raycache_gather")
        return buf

    def _enqueue(self):
        buf = self._bufs[self._slot]
        self._slot ^= 1
        st = self.stream
        if st is not None:
            st.wait_stream(torch.cuda.current_stream(self.device))     # the consumer of this buffer (2 steps ago) is done
            with torch.cuda.stream(st):
                idx = self._next_indices()
                self.gather(idx, buf)
                buf["n_host"].copy_(buf["n_valid"], non_blocking=True)
                buf["event"].record(st)
        else:
            idx = self._next_indices()
            self.gather(idx, buf)
            buf["n_host"].copy_(buf["n_valid"], non_blocking=True)
            buf["event"].record()
        return buf

    def next_batch(self):
        buf = self._pending if self._pending is not None else self._enqueue()
        buf["event"].synchronize()                       # only the loader's own small kernels, issued a step ago
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        m = int(buf["n_host"][0])
        out = {"rays": buf["rays"][:m], "rgbs": buf["rgbs"][:m], "ts": buf["ts"][:m], "label": buf["label"][:m], "n_valid": m}
        self._pending = self._enqueue() if self.stream is not None else None
        return out

    def __iter__(self):
        while True:
            yield self.next_batch()


def synthetic_cache(n_rays, n_images=64, origin=(0.0, 0.0, 0.0), radius=1.0, n_vocab=5000, seed=1, masked_fraction=0.0,
                    H=400, W=400, focal=400.0):
    """Synthetic posed-camera ray cache in the reference layout (host tensors): `n_images` pinhole cameras on a ring
    around the scene looking at its centre (datasets/ray_utils.py:5-52 ray construction), ts = image id, labels
    from {wall, building, sky, road} plus a `masked_fraction` of person / car rows for the filter."""
    from .renderer import LABEL_IDS

    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, n_images, (n_rays,), generator=g)
    pix = torch.randint(0, H * W, (n_rays,), generator=g)
    i, j = (pix % W).float(), (pix // W).float()
    d_cam = torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)
    ang = img.float() / n_images * 2 * np.pi
    cam = torch.stack([3.0 * torch.sin(ang), 0.3 * torch.cos(3 * ang), -3.0 * torch.cos(ang)], -1)   # ring of radius 3
    fwd = -cam / cam.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    up2 = torch.cross(right, fwd, dim=-1)
    d = d_cam[:, :1] * right + d_cam[:, 1:2] * up2 - d_cam[:, 2:3] * fwd        # camera -z looks along fwd
    d = d / d.norm(dim=-1, keepdim=True)
    org = torch.tensor(origin, dtype=torch.float32)
    o = cam * radius + org
    near = torch.full((n_rays, 1), 2.0 * radius)
    far = torch.full((n_rays, 1), 4.0 * radius)
    ts = (img % n_vocab).float().unsqueeze(1)
    base = torch.tensor([0.0, 1.0, 2.0, 6.0])[torch.randint(0, 4, (n_rays,), generator=g)]
    bad = torch.tensor([float(LABEL_IDS["person"]), float(LABEL_IDS["car"])])[torch.randint(0, 2, (n_rays,), generator=g)]
    label = torch.where(torch.rand(n_rays, generator=g) < masked_fraction, bad, base).unsqueeze(1)
    has = (torch.rand(n_rays, generator=g) < 0.2).float()
    depth = ((near + (far - near) * torch.rand(n_rays, 1, generator=g)).squeeze(1) * has).unsqueeze(1)
    weight = (2.0 * (1.0 - torch.rand(n_rays, generator=g)) * has).unsqueeze(1)
    rays = torch.cat([o, d, near, far, ts, label, depth, weight], 1).float().contiguous()
    rgbs = torch.rand(n_rays, 3, generator=g)
    return rays, rgbs
