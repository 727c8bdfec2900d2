"""numpy restatement of the octree half of the path (TEST INFRASTRUCTURE; parity UNPINNED).

The reference delegates these steps to NVIDIA Kaolin's SPC ops (fork git+https://github.com/Burningdust21/kaolin.git,
unpinned, environment.yaml:18), which is neither under /root/reference nor installable here.  This file restates
Kaolin's DOCUMENTED semantics as used by the reference's call sites:

  tools/prepare_data/generate_voxel.py:149-150  quantize_points + unbatched_points_to_octree
  tools/prepare_data/generate_voxel.py:173-178  scan_octrees + generate_points  (pyramid / prefix / point hierarchy)
  tools/prepare_data/generate_voxel.py:311-439  get_near_far over unbatched_raytrace(level, return_depth=True)

Encoding: octree = one byte per non-leaf node, breadth-first, Morton-sorted within a level (x is the most
significant bit of a Morton digit); bit j of a byte <=> child with digit j = (x&1)<<2 | (y&1)<<1 | (z&1) exists;
prefix = exclusive popcount sum (children of hierarchy node i start at 1 + prefix[i]); pyramid[0,l] = #nodes of
level l, pyramid[1,l] = first hierarchy index of level l.

A voxel is hit iff the fp32 slab test below passes for it and for all its ancestors; every step is a single
IEEE binary32 operation in the order written in csrc/octree.cu, so the CUDA tracer must agree bit for bit.
"""
import numpy as np

f32 = np.float32


def quantize_points(x, level):
    """kaolin.ops.spc.points.quantize_points: floor(clamp(2^L (x+1)/2, 0, 2^L - 1))."""
    res = 2 ** level
    q = np.floor(np.clip(res * (np.asarray(x, np.float64) + 1.0) / 2.0, 0, res - 1))
    return q.astype(np.int32)


def _morton(c, level):
    m = np.zeros(len(c), np.int64)
    for i in range(level):
        m |= ((c[:, 0] >> i) & 1).astype(np.int64) << (3 * i + 2)
        m |= ((c[:, 1] >> i) & 1).astype(np.int64) << (3 * i + 1)
        m |= ((c[:, 2] >> i) & 1).astype(np.int64) << (3 * i)
    return m


def build_octree(points_normalized, level):
    """unbatched_points_to_octree + scan_octrees + generate_points.
    Returns dict(octree u8[n_nonleaf], prefix i32[n_nonleaf], pyramid i32[2,level+2], points i16[n_total,3])."""
    q = quantize_points(points_normalized, level)
    levels = [None] * (level + 1)
    levels[level] = np.unique(q, axis=0)
    for l in range(level - 1, -1, -1):
        levels[l] = np.unique(levels[l + 1] >> 1, axis=0)
    for l in range(level + 1):
        levels[l] = levels[l][np.argsort(_morton(levels[l], l), kind="stable")]
    octree = []
    for l in range(level):
        parents, children = levels[l], levels[l + 1]
        key = {tuple(p): i for i, p in enumerate(parents)}
        bytes_l = np.zeros(len(parents), np.uint8)
        for ch in children:
            j = ((ch[0] & 1) << 2) | ((ch[1] & 1) << 1) | (ch[2] & 1)
            bytes_l[key[tuple(ch >> 1)]] |= np.uint8(1 << j)
        octree.append(bytes_l)
    octree = np.concatenate(octree) if octree else np.zeros(0, np.uint8)
    pop = np.array([bin(int(b)).count("1") for b in octree], np.int32)
    prefix = np.concatenate([[0], np.cumsum(pop)[:-1]]).astype(np.int32) if len(pop) else np.zeros(0, np.int32)
    pyramid = np.zeros((2, level + 2), np.int32)
    for l in range(level + 1):
        pyramid[0, l] = len(levels[l])
    pyramid[1, 1:] = np.cumsum(pyramid[0, :-1])
    points = np.concatenate(levels).astype(np.int16)
    return dict(octree=octree, prefix=prefix, pyramid=pyramid, points=points, levels=levels)


def _slab(o, d, coords, l):
    """fp32 slab test of rays (o,d [R,3]) against voxels coords [V,3] of level l -> (hit [R,V], depth [R,V])."""
    r = f32(1.0) / f32(2 ** l)
    c = (r * (2 * coords + 1).astype(f32) - f32(1.0)).astype(f32)              # [V,3]
    lo = ((c - r).astype(f32)[None] - o[:, None]).astype(f32)
    hi = ((c + r).astype(f32)[None] - o[:, None]).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo / d[:, None]).astype(f32)
        t2 = (hi / d[:, None]).astype(f32)
    tmin = np.maximum.reduce(np.minimum(t1, t2), axis=-1)
    tmax = np.minimum.reduce(np.maximum(t1, t2), axis=-1)
    hit = (tmax >= tmin) & (tmax >= 0)
    return hit, np.maximum(tmin, f32(0.0)).astype(f32)


def normalise_rays(rays_o, rays_d, scene_origin, scale):
    """generate_voxel.py:332-333,345 in fp32."""
    d = (np.asarray(rays_d, f32) + f32(1e-7)).astype(f32)
    o = (((np.asarray(rays_o, f32) + f32(1e-7)).astype(f32) - np.asarray(scene_origin, f32)).astype(f32) / f32(scale)).astype(f32)
    return o, d


def raytrace(tree, level, rays_o, rays_d, scene_origin, scale):
    """Hit list of unbatched_raytrace(level, return_depth=True, with_exit=False): (ray_index, point_index, depth),
    grouped by ray, front-to-back (ties: Morton / hierarchy index)."""
    o, d = normalise_rays(rays_o, rays_d, scene_origin, scale)
    leaves = tree["levels"][level]
    ok = np.ones((len(o), len(leaves)), bool)
    for l in range(0, level):
        anc = leaves >> (level - l)
        h, _ = _slab(o, d, anc, l)
        ok &= h
    h, depth = _slab(o, d, leaves, level)
    ok &= h
    base = int(tree["pyramid"][1, level])
    ri, li = np.nonzero(ok)
    dep = depth[ri, li]
    order = np.lexsort((li, dep, ri))
    return ri[order].astype(np.int32), (li[order] + base).astype(np.int32), dep[order].astype(f32)


def get_near_far(tree, level, rays_o, rays_d, scene_origin, scale):
    """get_near_far post-processing (generate_voxel.py:374-400,437-439): first hit depth, last hit ENTRY depth,
    invalidated when near <= 1e-4; returned multiplied by scale.  Also pid (hierarchy index, -1 = miss), count."""
    R = len(rays_o)
    ri, pi, dep = raytrace(tree, level, rays_o, rays_d, scene_origin, scale)
    near = np.zeros(R, f32)
    far = np.zeros(R, f32)
    pid = -np.ones(R, np.int32)
    count = np.bincount(ri, minlength=R).astype(np.int32)
    for r in range(R):
        sel = ri == r
        if sel.any():
            near[r], far[r], pid[r] = dep[sel][0], dep[sel][-1], pi[sel][0]
    bad = ~(near > f32(1e-4))
    near[bad] = 0
    far[bad] = 0
    pid[bad] = -1
    return (near * f32(scale)).astype(f32), (far * f32(scale)).astype(f32), pid, count


def sphere_shell_points(radius=0.5, voxel=0.05, seed=0, n=20000):
    """Synthetic surface point cloud: the shell | |x| - radius | < voxel."""
    rng = np.random.RandomState(seed)
    v = rng.randn(n, 3)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * (radius + (rng.rand(n, 1) * 2 - 1) * voxel)


# ---- K0 host math around the builder (tools/prepare_data/generate_voxel.py:27-38, 75-150) -------------------------
def expand_points(points, voxel_size):
    """generate_voxel.py:27-38: 3x3x3 dilation by voxel_size, unique rows."""
    from itertools import product
    grids = list(product(*zip([-1, -1, -1], [0, 0, 0], [1, 1, 1])))
    ex = np.concatenate([points + np.array(g) * voxel_size for g in grids], axis=0)
    return np.unique(ex, axis=0)


def gen_octree(scene_config, points, voxel_size, expand=1, radius=1.0, in_sfm=True):
    """generate_voxel.py:75-150 without file I/O (scene_config = the dict read from config.yaml).
    Returns (tree, scene_origin, scale, level, points_filtered)."""
    points = np.asarray(points, np.float64)
    if in_sfm:
        gt_to_sfm = np.linalg.inv(np.array(scene_config["sfm2gt"]))
        v1 = gt_to_sfm[:3, :3] @ np.array(scene_config["eval_bbx"][0]) + gt_to_sfm[:3, 3]
        v2 = gt_to_sfm[:3, :3] @ np.array(scene_config["eval_bbx"][1]) + gt_to_sfm[:3, 3]
        bbx_min, bbx_max = np.minimum(v1, v2), np.maximum(v1, v2)
    else:
        bbx_min, bbx_max = np.array(scene_config["eval_bbx"][0]), np.array(scene_config["eval_bbx"][1])
    dim = np.max(bbx_max - bbx_min)
    for _ in range(expand):
        points = expand_points(points, voxel_size)
    scene_origin = bbx_min + (bbx_max - bbx_min) / 2
    scale = dim / 2 * radius
    pn = (points - scene_origin) / scale
    mask = np.prod((pn > -1), axis=-1, dtype=bool) & np.prod((pn < 1), axis=-1, dtype=bool)
    pf = pn[mask]
    level = int(np.floor(np.log2(2 * scale / voxel_size)))
    return build_octree(pf, level), scene_origin, scale, level, pf


# ---- octree refresh (lightning_modules/neuconw_system.py:186-312) ---------------------------------------------------
def surface_selection(tree, octree_origin, octree_scale, octree_level, train_level, threshold, sdf_fn, scene_origin_sfm,
                      scene_radius_sfm):
    """neuconw_system.py:186-266 in the reference's dtypes: candidate voxels = every level-`octree_level` leaf split
    2^(train_level-octree_level) times per axis; kept where sdf_fn(xyz_training float32 [n,3]) <= threshold.
    Returns (sparse_pc_sfm float32 [m,3], train_voxel_size)."""
    import torch
    leaves = tree["levels"][octree_level]
    order = np.lexsort((leaves[:, 2], leaves[:, 1], leaves[:, 0]))          # torch.nonzero(dense) order
    sparse_ind = torch.from_numpy(leaves[order].astype(np.int64))
    sparse_num = sparse_ind.shape[0]
    up_times = 2 ** (train_level - octree_level)
    sparse_ind_up = sparse_ind.repeat_interleave(up_times ** 3, dim=0) * up_times
    k = torch.arange(0, up_times, 1)
    up_kernal = torch.stack(torch.meshgrid(k, k, k, indexing="ij"), dim=-1).reshape(-1, 3)
    sparse_ind_up = sparse_ind_up + up_kernal.repeat([sparse_num, 1])
    train_voxel_size = 2 / (2 ** train_level) * octree_scale
    origin32 = torch.as_tensor(np.asarray(octree_origin)).float()
    vol_origin = origin32 - octree_scale
    xyz_sfm = sparse_ind_up * train_voxel_size + vol_origin                   # float32
    xyz_training = (xyz_sfm - torch.as_tensor(np.asarray(scene_origin_sfm)).float()) / scene_radius_sfm
    sdf = np.asarray(sdf_fn(xyz_training.numpy()), np.float32).reshape(-1)
    return xyz_sfm.numpy()[sdf <= threshold], train_voxel_size


def octree_update(scene_config, tree, octree_origin, octree_scale, octree_level, train_level, threshold, sdf_fn,
                  scene_origin_sfm, scene_radius_sfm):
    """neuconw_system.py:268-312: surface_selection + gen_octree(expand=False)."""
    pc, tvs = surface_selection(tree, octree_origin, octree_scale, octree_level, train_level, threshold, sdf_fn,
                                scene_origin_sfm, scene_radius_sfm)
    new_tree, origin, scale, level, _ = gen_octree(scene_config, pc, tvs, expand=0)
    return new_tree, origin, scale, level, tvs, pc
