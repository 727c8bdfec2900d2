"""Synthetic posed-camera ray batches in the reference's ray-cache layout (host utility).

Mirrors datasets/ray_utils.py:5-52 (pinhole camera, no +0.5 pixel offset, camera-space dirs
[(i-cx)/fx, -(j-cy)/fy, -1] rotated by c2w, L2-normalised) and the cache row layout of
datasets/phototourism.py:611-623: rays [n,10] = o3, d3, near, far, depth, depth_weight; ts; label.
"""
import torch


def make_ray_batch(n_rays, origin=(0.0, 0.0, 0.0), radius=1.0, n_vocab=5000, seed=1, H=400, W=400, focal=400.0,
                   device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    pix = torch.randint(0, H * W, (n_rays,), generator=g)
    i = (pix % W).float()
    j = (pix // W).float()
    d_cam = torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)
    rot = torch.diag(torch.tensor([-1.0, 1.0, -1.0]))  # camera looks along world +z
    d = d_cam @ rot.T
    d = d / d.norm(dim=-1, keepdim=True)
    org = torch.tensor(origin, dtype=torch.float32)
    o = (torch.tensor([0.0, 0.0, -3.0]) * radius + org).expand(n_rays, 3)
    near = torch.full((n_rays, 1), 2.0 * radius)
    far = torch.full((n_rays, 1), 4.0 * radius)
    has = (torch.rand(n_rays, generator=g) < 0.2).float()
    dgt = (near + (far - near) * torch.rand(n_rays, 1, generator=g)).squeeze(1) * has
    dw = 2.0 * (1.0 - torch.rand(n_rays, generator=g)) * has
    rays = torch.cat([o, d, near, far, dgt[:, None], dw[:, None]], 1).float().contiguous()
    ts = torch.randint(0, n_vocab, (n_rays,), generator=g)
    label = torch.tensor([0.0, 1.0, 2.0, 6.0])[torch.randint(0, 4, (n_rays,), generator=g)]
    rgbs = torch.rand(n_rays, 3, generator=g)
    out = dict(rays=rays, ts=ts, label=label, rgbs=rgbs)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


def sphere_shell_points(radius=0.5, thickness=0.05, n=20000, seed=0):
    """Synthetic SfM point cloud: points in the shell | |x| - radius | < thickness (stands in for COLMAP's
    points3D.bin of a real scene)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=-1, keepdim=True)
    return v * (radius + (torch.rand(n, 1, generator=g, dtype=torch.float64) * 2 - 1) * thickness)


def install_synthetic_scene(renderer, radius=1.0, origin=(0.0, 0.0, 0.0), voxel_size=0.1, n_points=20000, seed=0):
    """What NeuconWSystem / get_octree read from <ROOT_DIR>/config.yaml and points3D.bin (neuconw_system.py:65-67,
    generate_voxel.py:41-73), synthesised: scene frame, eval bounding box, SfM point cloud."""
    import numpy as np

    r = float(radius)
    renderer.scene_config = {"sfm2gt": np.eye(4).tolist(),
                             "eval_bbx": [[origin[0] - r, origin[1] - r, origin[2] - r], [origin[0] + r, origin[1] + r, origin[2] + r]]}
    renderer.sfm_points = (sphere_shell_points(0.5 * r, 0.03 * r, n_points, seed) + torch.tensor(origin, dtype=torch.float64)).numpy()
    renderer.voxel_size = voxel_size * r
    return renderer
