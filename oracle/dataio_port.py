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
    n_items = len(items)
    items_permute = np.random.RandomState(seed).permutation(items)
    if n_items % world_size == 0:
        padded = items_permute
    else:
        padding = np.random.RandomState(seed).choice(items, world_size - (n_items % world_size), replace=True)
        padded = np.concatenate([items_permute, padding])
    per = len(padded) // world_size
    return padded[per * rank: per * (rank + 1)]


def getitem_batch(all_rays, all_rgbs, index):
    """all_rays [n,12], all_rgbs [n,3] torch CPU; index int64 [B] -> dict like the collated DataLoader batch."""
    rows = all_rays[index]
    return {"rays": torch.cat((rows[:, :8], rows[:, 10:13]), dim=-1), "ts": rows[:, 8].long(), "rgbs": all_rgbs[index],
            "semantics": rows[:, 9]}


def filter_batch(batch, ray_mask_list=("person", "car", "bicycle", "minibike")):
    ts, label = batch["ts"], batch["semantics"]
    ray_mask = torch.ones_like(ts, dtype=torch.bool)
    for name in ray_mask_list or ():
        ray_mask[LABEL_IDS[name] == label] = False
    return {"rays": batch["rays"][ray_mask, :], "ts": ts[ray_mask], "rgbs": batch["rgbs"][ray_mask], "label": label[ray_mask]}


def dense_lattice(dim, origin=(0.0, 0.0, 0.0), radius=1.0):
    so = np.array(origin)
    x = torch.linspace(so[0] - radius, so[0] + radius, dim)
    y = torch.linspace(so[1] - radius, so[1] + radius, dim)
    z = torch.linspace(so[2] - radius, so[2] + radius, dim)
    return torch.stack(torch.meshgrid(x, y, z, indexing="ij"), dim=-1).reshape(-1, 3)


def sparse_lattice(sparse_ind, up_times, voxel_size, vol_origin, scene_origin, scene_radius):
    """sparse_ind int64 [m,3] (torch.nonzero order); returns (xyz_sfm, xyz_training) float32 as the reference computes them."""
    sparse_num = sparse_ind.shape[0]
    up = sparse_ind.repeat_interleave(up_times ** 3, dim=0) * up_times
    k = torch.arange(0, up_times, 1)
    up_kernal = torch.stack(torch.meshgrid(k, k, k, indexing="ij"), dim=-1).reshape(-1, 3)
    up = up + up_kernal.repeat([sparse_num, 1])
    xyz_sfm = up * voxel_size + vol_origin
    return xyz_sfm, (xyz_sfm - scene_origin) / scene_radius


def local_range(n, world_size, rank):
    per = n // world_size if n % world_size == 0 else n // world_size + 1
    return rank * per, min(n, (rank + 1) * per), per
