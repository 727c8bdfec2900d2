"""Device-resident SDF-grid evaluation for mesh extraction (BASELINE config 5).

Mirrors the SDF half of utils/visualization.py::extract_mesh (lines 36-107) and tools/extract_mesh.py::gen_grid_spc
(lines 60-102).  The reference builds the query lattice on the CPU, ships every chunk to the GPU, copies every SDF
chunk back (`.cpu()` per chunk), all-gathers and scatters into a dense volume on the host.  Here the lattice points are
GENERATED on the device chunk by chunk (nrw_grid_points_dense / nrw_grid_points_sparse) straight into the SDF query
(nrw_sdf_query), ranks take the contiguous slices of get_local_split (utils/visualization.py:27-35) and one
all_gather joins them; the dense volume, the validity mask and the scatter stay on the GPU.  Marching cubes
itself (skimage.measure.marching_cubes) and trimesh export are host-side post-processing and out of scope:
`sdf_volume` / `sparse_sdf_volume` return exactly the arrays the reference hands to marching_cubes."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import NrwError, check, ptr, stream_ptr


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _local_range(n, world, rank):
    """get_local_split (utils/visualization.py:27-35): pad to a multiple of world, contiguous slice per rank.
    Returns (start, stop_unpadded, slice_length)."""
    per = n // world if n % world == 0 else n // world + 1
    a = rank * per
    return a, min(n, a + per), per


def _gather_slices(local, per, n, world):
    """all_gather of equal-length slices (utils/visualization.py:81-88) -> first n entries, on every rank."""
    if world == 1:
        return local[:n]
    if local.shape[0] < per:                  # this rank's slice reaches into the zero padding
        local = torch.cat([local, torch.zeros(per - local.shape[0], dtype=local.dtype, device=local.device)])
    parts = [torch.empty(per, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())                   # NCCL over NVLink in training; the reference's collective
    return torch.cat(parts, 0)[:n]


def sdf_volume(renderer, dim, origin=(0.0, 0.0, 0.0), radius=1.0, chunk=1 << 20):
    """Dense branch of extract_mesh (sparse_data=None): SDF on the dim^3 lattice of
    torch.linspace(origin[c]-radius, origin[c]+radius, dim), returned as a float32 CUDA tensor [dim,dim,dim]
    (= `sdf.reshape((dim,dim,dim))` of utils/visualization.py:96), plus (vol_origin, voxel_size)."""
    L = _lib.lib()
    eng = renderer.engine
    dev = next(renderer.neuconw.parameters()).device
    if dev.type != "cuda":
        raise NrwError("sdf_volume: the networks must live on a CUDA device")
    o64 = np.array(origin, dtype=np.float64)
    lo, hi = (o64 - radius).astype(np.float32), (o64 + radius).astype(np.float32)
    n = int(dim) ** 3
    world, rank = _world()
    a, b, per = _local_range(n, world, rank)
    local = torch.empty(max(b - a, 0), dtype=torch.float32, device=dev)
    pts = torch.empty(min(chunk, max(b - a, 1)), 3, dtype=torch.float32, device=dev)
    eng.ensure(dev, 1, 2, 0, chunk_hint=pts.shape[0])
    eng.pack(dev)
    with torch.no_grad():
        for i in range(a, b, pts.shape[0]):
            m = min(pts.shape[0], b - i)
            check(L.nrw_grid_points_dense(int(dim), _f3(lo), _f3(hi), i, m, ptr(pts), stream_ptr()), "nrw_grid_points_dense")
            check(L.nrw_sdf_query(eng.ctx, ptr(pts), m, C.c_void_p(local.data_ptr() + (i - a) * 4), stream_ptr()), "nrw_sdf_query")
    sdf = _gather_slices(local, per, n, world)
    return sdf.reshape(dim, dim, dim), (o64 - radius), 2 * radius / (dim - 1)


def gen_grid_spc(renderer, eval_level, device=0):
    """tools/extract_mesh.py:60-102 without materialising the candidate list: returns the description of the
    up-sampled sparse lattice {leaves int16 [m,3] (lexicographic), up_times, voxel_size, dim, vol_origin}."""
    if renderer.octree_data is None:
        renderer.octree_data = renderer.get_octree(device)
    od = renderer.octree_data
    spc = od["spc_data"]
    L0 = int(od["level"])
    pyr = spc["pyramid"]
    leaves = spc["points"][int(pyr[1, L0]):int(pyr[1, L0 + 1])].to(torch.int64)
    key = (leaves[:, 0] * (1 << L0) + leaves[:, 1]) * (1 << L0) + leaves[:, 2]
    leaves = leaves[torch.argsort(key)].to(torch.int16).contiguous()          # torch.nonzero(dense) order
    up_times = 2 ** (int(eval_level) - L0)
    if up_times < 1:
        raise NrwError(f"gen_grid_spc: eval_level {eval_level} below the octree level {L0}")
    scale = od["scale"]
    return {"leaves": leaves, "up_times": up_times, "voxel_size": 2 / (2 ** int(eval_level)) * scale,
            "dim": int((2 ** L0) * up_times), "vol_origin": od["scene_origin"].float().cpu() - scale}


def sparse_candidates(renderer, grid, chunk=1 << 20, threshold=None, sdf_fn=None, want_sdf=True):
    """SDF of every candidate of an up-sampled sparse lattice (gen_grid_spc / surface_selection), chunked:
    points generated on the device, SDF through nrw_sdf_query (or `sdf_fn(xyz_training)` in tests), ranks split and
    all-gather as the reference does (neuconw_system.py:236-258).  With `threshold` the stable compaction
    xyz_sfm[sdf <= threshold] (neuconw_system.py:259) runs chunk by chunk on the device as well.
    Returns dict(sdf [n] or None, kept_xyz_sfm [m,3] or None, n)."""
    L = _lib.lib()
    eng = renderer.engine
    leaves = grid["leaves"]
    dev = leaves.device
    up = int(grid["up_times"])
    n = int(leaves.shape[0]) * up ** 3
    voxel = float(np.float32(grid["voxel_size"]))
    vol_origin = [float(x) for x in grid["vol_origin"]]
    scene_origin = [float(x) for x in torch.as_tensor(renderer.origin).float().cpu()]
    radius = float(renderer.radius)
    world, rank = _world()
    a, b, per = _local_range(n, world, rank)
    local = torch.empty(max(b - a, 0), dtype=torch.float32, device=dev)
    c = min(chunk, max(b - a, 1))
    xt = torch.empty(c, 3, dtype=torch.float32, device=dev)
    if sdf_fn is None:
        eng.ensure(dev, 1, 2, 0, chunk_hint=c)
        eng.pack(dev)
    with torch.no_grad():
        for i in range(a, b, c):
            m = min(c, b - i)
            check(L.nrw_grid_points_sparse(ptr(leaves), leaves.shape[0], up, voxel, _f3(vol_origin), _f3(scene_origin), radius,
                                           i, m, None, ptr(xt), stream_ptr()), "nrw_grid_points_sparse")
            if sdf_fn is None:
                check(L.nrw_sdf_query(eng.ctx, ptr(xt), m, C.c_void_p(local.data_ptr() + (i - a) * 4), stream_ptr()), "nrw_sdf_query")
            else:
                local[i - a:i - a + m] = sdf_fn(xt[:m]).reshape(-1)
    sdf = _gather_slices(local, per, n, world)
    kept = None
    if threshold is not None:
        # capacity: every candidate for small lattices, an exact count first for large ones
        total = n if n <= (1 << 24) else int((sdf <= threshold).sum())
        cap = torch.empty(max(total, 1), 3, dtype=torch.float32, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        xs = torch.empty(c, 3, dtype=torch.float32, device=dev)
        sb = L.nrw_compact_scratch_bytes(c)
        scratch = torch.empty(sb + 256, dtype=torch.uint8, device=dev)
        sp = (scratch.data_ptr() + 255) // 256 * 256
        for i in range(0, n, c):
            m = min(c, n - i)
            check(L.nrw_grid_points_sparse(ptr(leaves), leaves.shape[0], up, voxel, _f3(vol_origin), _f3(scene_origin), radius,
                                           i, m, ptr(xs), ptr(xt), stream_ptr()), "nrw_grid_points_sparse")
            check(L.nrw_threshold_compact(C.c_void_p(sdf.data_ptr() + i * 4), ptr(xs), m, float(threshold), ptr(cap), ptr(count),
                                          C.c_void_p(sp), stream_ptr()), "nrw_threshold_compact")
        kept = cap[:int(count)]
    return {"sdf": sdf if want_sdf else None, "kept_xyz_sfm": kept, "n": n}


def sparse_sdf_volume(renderer, grid, chunk=1 << 20, sdf_fn=None):
    """Sparse branch of extract_mesh (utils/visualization.py:53-66,98-116): returns (sdf_dense [dim]^3 float32 CUDA, ones
    outside the lattice; mask_dense bool CUDA - a cell is valid iff all 8 of its corners are lattice points)."""
    res = sparse_candidates(renderer, grid, chunk=chunk, sdf_fn=sdf_fn)
    leaves, up, dim = grid["leaves"].to(torch.int64), int(grid["up_times"]), int(grid["dim"])
    dev = leaves.device
    k = torch.arange(up, device=dev)
    kern = torch.stack(torch.meshgrid(k, k, k, indexing="ij"), -1).reshape(-1, 3)
    ind = (leaves[:, None, :] * up + kern[None, :, :]).reshape(-1, 3)          # ind = round((xyz - vol_origin)/voxel)
    sdf_dense = torch.ones(dim, dim, dim, dtype=torch.float32, device=dev)
    sdf_dense[ind[:, 0], ind[:, 1], ind[:, 2]] = res["sdf"]
    m = torch.zeros(dim, dim, dim, dtype=torch.bool, device=dev)
    m[ind[:, 0], ind[:, 1], ind[:, 2]] = True
    # a marching-cubes cell (i, j, k) uses the corners (i - a, j - b, k - c), a, b, c in {0, 1} (wrap-around as torch.roll)
    valid = m.clone()
    for shift in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)):
        valid &= torch.roll(m, shifts=shift, dims=(0, 1, 2))
    m = valid
    return sdf_dense, m
