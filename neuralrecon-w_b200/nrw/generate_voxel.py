"""Kaolin-free stand-ins, WITH THE REFERENCE'S SIGNATURES, for the names lightning_modules/neuconw_system.py:26-30 and
tools/extract_mesh.py import from tools/prepare_data/generate_voxel.py:

    convert_to_dense(octree, level)                                     generate_voxel.py:181-186
    gen_octree(recontruct_path, points, voxel_size, device, visualize, expand, radius)   generate_voxel.py:75-171
    octree_to_spc(octree)                                               generate_voxel.py:173-178

so that NeuconWSystem.surface_selection / octree_update (neuconw_system.py:186-312) run unmodified on top of the CUDA
octree builder (csrc/octree_build.cu):

    -from tools.prepare_data.generate_voxel import convert_to_dense, gen_octree, octree_to_spc
    +from nrw.generate_voxel import convert_to_dense, gen_octree, octree_to_spc

`octree` is the breadth-first child-mask byte tensor Kaolin calls an octree; `octree_to_spc` decodes it level by level
on the GPU (a few torch ops per level, not on the hot path) into the same (points int16 [n,3], pyramid int32 [2,L+2],
prefix int32 [n_nonleaf]) triple `spc.scan_octrees` + `spc.generate_points` return (semantics: SURVEY.md 8c)."""
import os

import numpy as np
import torch

from ._lib import NrwError
from . import octree as _oct


def octree_to_spc(octree):
    if octree.dtype != torch.uint8 or octree.dim() != 1:
        raise NrwError("octree_to_spc: expected the uint8 child-mask byte tensor of an octree")
    dev = octree.device
    n_bytes = octree.shape[0]
    counts = torch.zeros(n_bytes, dtype=torch.int32, device=dev)
    for j in range(8):
        counts += ((octree >> j) & 1).to(torch.int32)
    prefix = (torch.cumsum(counts, 0) - counts).to(torch.int32)                     # exclusive sum of the popcounts
    coords = torch.zeros(1, 3, dtype=torch.int64, device=dev)
    levels, sizes, off = [coords], [1], 0
    shifts = torch.arange(8, device=dev)
    while off < n_bytes:
        n = coords.shape[0]
        if off + n > n_bytes:
            raise NrwError("octree_to_spc: truncated octree")
        bits = (octree[off:off + n, None].to(torch.int64) >> shifts[None, :]) & 1     # [n, 8]
        node, child = torch.nonzero(bits, as_tuple=True)                               # node-major, child ascending = Morton order
        delta = torch.stack([(child >> 2) & 1, (child >> 1) & 1, child & 1], -1)      # x most significant
        coords = coords[node] * 2 + delta
        off += n
        levels.append(coords)
        sizes.append(coords.shape[0])
    L = len(sizes) - 1
    pyramid = torch.zeros(2, L + 2, dtype=torch.int32)
    pyramid[0, :L + 1] = torch.tensor(sizes, dtype=torch.int32)
    pyramid[1, 1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0).to(torch.int32)
    points = torch.cat(levels, 0).to(torch.int16)
    return points, pyramid, prefix


def convert_to_dense(octree, level):
    points, pyramid, _ = octree_to_spc(octree)
    return _oct.convert_to_dense({"points": points, "pyramid": pyramid}, level)


def gen_octree(recontruct_path, points, voxel_size, device=0, visualize=False, expand=1, radius=1.0, in_sfm=True):
    if visualize:
        raise NrwError("gen_octree(visualize=True) writes Open3D debug point clouds in the reference; not implemented")
    import yaml

    with open(os.path.join(recontruct_path, "config.yaml"), "r") as f:
        scene_config = yaml.load(f, Loader=yaml.FullLoader)
    pts = points if torch.is_tensor(points) else torch.from_numpy(np.asarray(points))
    tree, scene_origin, scale, level = _oct.gen_octree(scene_config, pts, voxel_size, device=device, expand=int(expand),
                                                       radius=radius, in_sfm=in_sfm)
    return tree["octree"], scene_origin, scale, level
