"""numpy / torch-CPU restatements of the data movers either side of the hot path (TEST INFRASTRUCTURE).

  local_split          DataModule._get_local_split                      datasets/data.py:83-100
  getitem_batch        PhototourismDataset.__getitem__ (train, semantics) datasets/phototourism.py:709-724,
                       stacked over an index vector as torch's default collate does
  filter_batch         RAY_MASK_LIST black list of training_step        lightning_modules/neuconw_system.py:345-355
  dense_lattice        extract_mesh, sparse_data=None                    utils/visualization.py:42-52
  sparse_lattice       gen_grid_spc / surface_selection up-sampling      tools/extract_mesh.py:73-95, neuconw_system.py:213-234
  local_range          get_local_split                                   utils/visualization.py:27-35

Pinned against the unmodified reference where its code is importable without Kaolin / COLMAP data
(tests/test_dataio_oracle.py: _get_local_split, __getitem__, get_local_split)."""
import numpy as np
import torch

LABEL_IDS = {"sky": 2, "road": 6, "person": 12, "car": 20, "minibike": 116, "bicycle": 127}   # datasets/mask_utils.py


def local_split(items, world_size, rank, seed=6):
    """Which cache splits a rank loads: a seeded permutation of the split names, topped up to a multiple of the world size
    by a seeded draw WITH replacement from the original list, then cut into equal contiguous runs."""
    order = np.random.RandomState(seed).permutation(items)
    shortfall = (-len(items)) % world_size
    if shortfall:
        order = np.concatenate([order, np.random.RandomState(seed).choice(items, shortfall, replace=True)])
    run = len(order) // world_size
    return order[rank * run:(rank + 1) * run]


def getitem_batch(all_rays, all_rgbs, index):
    """Cache rows -> collated training batch.  all_rays [n,12] = o3, d3, near, far, ts, label, depth, weight."""
    picked = all_rays[index]
    geometry, extras = picked[:, 0:8], picked[:, 10:13]          # the reference slices 10:13 of a 12-wide row -> 2 columns
    return {"rays": torch.cat((geometry, extras), dim=-1), "ts": picked[:, 8].long(), "rgbs": all_rgbs[index],
            "semantics": picked[:, 9]}


def filter_batch(batch, ray_mask_list=("person", "car", "bicycle", "minibike")):
    """Drop the rays whose semantic label is on the black list; order of the survivors is kept."""
    label = batch["semantics"]
    keep = torch.ones(label.shape[0], dtype=torch.bool)
    for name in ray_mask_list or ():
        keep &= ~(label == LABEL_IDS[name])
    return {"rays": batch["rays"][keep], "ts": batch["ts"][keep], "rgbs": batch["rgbs"][keep], "label": label[keep]}


def dense_lattice(dim, origin=(0.0, 0.0, 0.0), radius=1.0):
    """dim^3 query lattice, x slowest / z fastest, each axis a float32 torch.linspace over [c - radius, c + radius]."""
    centre = np.array(origin)
    axes = [torch.linspace(centre[a] - radius, centre[a] + radius, dim) for a in range(3)]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, 3)


def sparse_lattice(sparse_ind, up_times, voxel_size, vol_origin, scene_origin, scene_radius):
    """Occupied coarse cells (int64 [m,3], torch.nonzero order), each refined into up_times^3 fine cells (cell-major, then
    the refinement offsets in ij-meshgrid order).  Returns (xyz_sfm, xyz_training) with the reference's dtype path:
    int64 index * python float -> float32, + float32 origin, then (x - scene_origin) / scene_radius."""
    step = torch.arange(0, up_times, 1)
    offsets = torch.stack(torch.meshgrid(step, step, step, indexing="ij"), dim=-1).reshape(-1, 3)
    fine = sparse_ind.repeat_interleave(up_times ** 3, dim=0) * up_times + offsets.repeat([sparse_ind.shape[0], 1])
    xyz_sfm = fine * voxel_size + vol_origin
    return xyz_sfm, (xyz_sfm - scene_origin) / scene_radius


def local_range(n, world_size, rank):
    """Row range of one rank when n rows are zero-padded up to a multiple of world_size and cut evenly:
    (first row, one past the last REAL row, rows per rank)."""
    per = -(-n // world_size)
    return rank * per, min(n, (rank + 1) * per), per
