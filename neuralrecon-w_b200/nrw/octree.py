"""K0 host mirror: the octree half of tools/prepare_data/generate_voxel.py on the CUDA builder (csrc/octree_build.cu).

    expand_points   generate_voxel.py:27-38    3x3x3 dilation + unique rows
    gen_octree      generate_voxel.py:75-150   bbox -> normalise -> filter -> level -> octree   (scene config passed in, no file I/O)
    octree_to_spc   generate_voxel.py:173-178  (the builder returns points / pyramid / prefix together with the bytes)
    convert_to_dense generate_voxel.py:181-186 dense occupancy of one level

Reading COLMAP's points3D.bin / config.yaml (gen_octree_from_sfm, generate_voxel.py:41-72) is data loading and stays
with the caller.  No CPU path: inputs must be CUDA tensors.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import NrwError, check, ptr, stream_ptr


def build_octree(points_normalized, level):
    """quantize_points + unbatched_points_to_octree + scan_octrees + generate_points.
    points_normalized: CUDA float32/float64 [N,3] inside (-1,1).  Returns dict(octree u8[n_nonleaf], prefix i32[n_nonleaf],
    pyramid i32[2,level+2] (host, as Kaolin returns it), points i16[n_total,3])."""
    if not points_normalized.is_cuda:
        raise NrwError("build_octree: points must be a CUDA tensor (no CPU path)")
    if points_normalized.dtype not in (torch.float32, torch.float64):
        raise NrwError(f"build_octree: unsupported dtype {points_normalized.dtype}")
    level = int(level)
    pts = points_normalized.detach().reshape(-1, 3).contiguous()
    n, dev = pts.shape[0], pts.device
    L = _lib.lib()
    cap_nl, cap_t = max(1, n * level), max(1, n * (level + 1))
    octree = torch.empty(cap_nl, dtype=torch.uint8, device=dev)
    prefix = torch.empty(cap_nl, dtype=torch.int32, device=dev)
    pyramid = torch.empty((2, level + 2), dtype=torch.int32, device=dev)
    pout = torch.empty((cap_t, 3), dtype=torch.int16, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    nb = int(L.nrw_octree_build_scratch_bytes(n, level, cap_nl))
    scratch = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    sp = (scratch.data_ptr() + 255) // 256 * 256
    check(L.nrw_octree_build(ptr(pts), 1 if pts.dtype == torch.float64 else 0, n, level, ptr(octree), ptr(prefix), ptr(pyramid),
                             ptr(pout), cap_nl, cap_t, ptr(counts), C.c_void_p(sp), stream_ptr()), "nrw_octree_build")
    n_nl, n_t = (int(v) for v in counts.cpu())      # the one host read of the build (sizes of the result tensors)
    if n_nl > cap_nl or n_t > cap_t:
        raise NrwError(f"build_octree: capacity exceeded ({n_nl}>{cap_nl} or {n_t}>{cap_t})")
    return {"octree": octree[:n_nl].clone(), "prefix": prefix[:n_nl].clone(), "pyramid": pyramid.cpu(),
            "points": pout[:n_t].clone()}


def expand_points(points, voxel_size):
    """generate_voxel.py:27-38 (the offsets are added in the input dtype, duplicates removed)."""
    grid = torch.tensor([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=points.dtype,
                        device=points.device)
    ex = (points[None, :, :] + grid[:, None, :] * voxel_size).reshape(-1, 3)
    return torch.unique(ex, dim=0)


def scene_bbox(scene_config, in_sfm=True):
    """generate_voxel.py:91-105 -> (bbx_min, bbx_max) float64 numpy."""
    if in_sfm:
        gt_to_sfm = np.linalg.inv(np.array(scene_config["sfm2gt"], dtype=np.float64))
        v1 = gt_to_sfm[:3, :3] @ np.array(scene_config["eval_bbx"][0], dtype=np.float64) + gt_to_sfm[:3, 3]
        v2 = gt_to_sfm[:3, :3] @ np.array(scene_config["eval_bbx"][1], dtype=np.float64) + gt_to_sfm[:3, 3]
        return np.minimum(v1, v2), np.maximum(v1, v2)
    return np.array(scene_config["eval_bbx"][0], dtype=np.float64), np.array(scene_config["eval_bbx"][1], dtype=np.float64)


def gen_octree(scene_config, points, voxel_size, device=0, expand=1, radius=1.0, in_sfm=True):
    """generate_voxel.py:75-150 with the scene config dict passed in.  points: [P,3] (numpy or tensor, SfM frame).
    Returns (octree_data, scene_origin, scale, level) where octree_data carries octree AND the spc tensors."""
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    pts = torch.as_tensor(points, dtype=torch.float64).to(dev)
    bbx_min, bbx_max = scene_bbox(scene_config, in_sfm)
    dim = float(np.max(bbx_max - bbx_min))
    for _ in range(int(expand)):
        pts = expand_points(pts, voxel_size)
    scene_origin = bbx_min + (bbx_max - bbx_min) / 2
    scale = dim / 2 * radius
    pn = (pts - torch.as_tensor(scene_origin, device=dev)) / scale
    mask = (pn > -1).all(dim=-1) & (pn < 1).all(dim=-1)
    pn = pn[mask]
    level = int(np.floor(np.log2(2 * scale / voxel_size)))
    tree = build_octree(pn, level)
    return tree, scene_origin, scale, level


def make_octree_data(scene_config, points, voxel_size, device=0, expand=1, radius=1.0, in_sfm=True):
    """The dict NeuconWRenderer.get_octree returns (renderer.py:137-155) / octree_update installs
    (neuconw_system.py:292-305)."""
    tree, scene_origin, scale, level = gen_octree(scene_config, points, voxel_size, device, expand, radius, in_sfm)
    dev = tree["octree"].device
    return {"octree": tree["octree"], "scene_origin": torch.from_numpy(np.asarray(scene_origin)).to(dev), "scale": scale,
            "level": level, "voxel_size": voxel_size,
            "spc_data": {"points": tree["points"], "pyramid": tree["pyramid"], "prefix": tree["prefix"]}}


def convert_to_dense(tree, level):
    """generate_voxel.py:181-186: dense [2^L]^3 float occupancy of `level` (torch scatter; not on the hot path)."""
    pyr = tree["pyramid"]
    a, b = int(pyr[1, level]), int(pyr[1, level + 1])
    p = tree["points"][a:b].long()
    res = 2 ** level
    dense = torch.zeros((res, res, res), dtype=torch.float32, device=p.device)
    dense[p[:, 0], p[:, 1], p[:, 2]] = 1.0
    return dense


def level_for_voxel(scale, voxel_size):
    """generate_voxel.py:146"""
    return int(math.floor(math.log2(2 * scale / voxel_size)))


# ---- octree refresh (lightning_modules/neuconw_system.py:186-312), device resident -------------------------------
def surface_selection(renderer, train_level, threshold, device=0, chunk=1 << 20, sdf_fn=None):
    """neuconw_system.py:186-266.  The reference densifies the coarse octree, moves it to the CPU, builds the candidate
    list with numpy-style torch ops, ships chunks to the GPU and every SDF chunk back; here the level-L leaves are taken
    straight from the point hierarchy (lexicographic order = torch.nonzero(dense)), candidate points are GENERATED on the
    GPU chunk by chunk in the same dtypes (nrw_grid_points_sparse: int64 index * python float -> float32), the SDF runs
    through nrw_sdf_query, ranks evaluate the contiguous slices of get_local_split and all-gather them
    (neuconw_system.py:236-258), and `xyz_sfm[sdf <= threshold]` is a stable device-side compaction.
    `sdf_fn(xyz_training)` overrides the network (tests).  Returns (sparse_pc_sfm float32 CUDA [m,3], train_voxel_size)."""
    from .mesh import gen_grid_spc, sparse_candidates

    grid = gen_grid_spc(renderer, train_level, device)
    res = sparse_candidates(renderer, grid, chunk=chunk, threshold=threshold, sdf_fn=sdf_fn, want_sdf=False)
    return res["kept_xyz_sfm"], grid["voxel_size"]


def octree_update(renderer, train_level, threshold, device=0, chunk=1 << 20, sdf_fn=None):
    """neuconw_system.py:268-312: installs renderer.fine_octree_data built from the current SDF (needs
    renderer.scene_config, as get_octree does).  With torch.distributed initialised the candidate SDFs are split across
    ranks and all-gathered exactly as the reference does (neuconw_system.py:236-258); every rank then builds the same
    octree."""
    cfg = getattr(renderer, "scene_config", None)
    if cfg is None:
        raise NrwError("octree_update: set renderer.scene_config (sfm2gt, eval_bbx of the scene's config.yaml)")
    renderer.fine_octree_data = None
    pc, tvs = surface_selection(renderer, train_level, threshold, device, chunk, sdf_fn)
    data = make_octree_data(cfg, pc, tvs, device=pc.device, expand=0)
    data["voxel_size"] = tvs
    renderer.fine_octree_data = data
    return data
