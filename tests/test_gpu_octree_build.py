"""K0: the CUDA octree builder against the numpy restatement of Kaolin's SPC construction
(oracle/octree_port.py; parity with Kaolin itself is UNPINNED - Kaolin is absent).  Integer work:
every output must be bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import octree_port as op
import util_nrw  # noqa: F401  (puts the package on sys.path)

pytestmark = pytest.mark.gpu


def nrw_pkg():
    import nrw
    import nrw.octree  # noqa: F401
    return nrw


def _cmp(tree, ref):
    assert np.array_equal(tree["pyramid"].numpy(), ref["pyramid"])
    assert np.array_equal(tree["octree"].cpu().numpy(), ref["octree"])
    assert np.array_equal(tree["prefix"].cpu().numpy(), ref["prefix"])
    assert np.array_equal(tree["points"].cpu().numpy(), ref["points"])


@pytest.mark.parametrize("level,n,dtype", [(1, 50, torch.float64), (3, 400, torch.float64), (5, 6000, torch.float32),
                                           (7, 20000, torch.float64), (9, 30000, torch.float32)])
def test_build_matches_restatement(level, n, dtype):
    nrw = nrw_pkg()
    pts = op.sphere_shell_points(0.6, 0.04, n=n, seed=level)
    t = torch.from_numpy(pts).to(dtype).cuda()
    tree = nrw.octree.build_octree(t, level)
    ref = op.build_octree(t.cpu().numpy(), level)       # the port promotes to float64 exactly as the kernel does
    _cmp(tree, ref)


def test_build_edge_cases():
    nrw = nrw_pkg()
    # duplicates, exact cell boundaries, clamped out-of-range values, the two cube corners
    pts = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], [0.999999, -0.999999, 0.5],
                    [0.25, 0.25, 0.25], [0.25, 0.25, 0.25 - 1e-12], [-3.0, 7.0, 0.1], [0.5, -0.5, -0.25]], np.float64)
    for level in (1, 2, 6, 15):
        _cmp(nrw.octree.build_octree(torch.from_numpy(pts).cuda(), level), op.build_octree(pts, level))
    one = np.array([[0.3, -0.2, 0.9]])
    _cmp(nrw.octree.build_octree(torch.from_numpy(one).cuda(), 8), op.build_octree(one, 8))
    empty = nrw.octree.build_octree(torch.zeros((0, 3), dtype=torch.float64, device="cuda"), 4)
    assert empty["octree"].numel() == 0 and empty["points"].numel() == 0 and int(empty["pyramid"].sum()) == 0
    with pytest.raises(nrw.NrwError):
        nrw.octree.build_octree(torch.zeros((4, 3), dtype=torch.float64, device="cuda"), 16)
    with pytest.raises(nrw.NrwError):
        nrw.octree.build_octree(torch.zeros((4, 3), dtype=torch.float64), 4)      # CPU tensor: no fallback


def test_gen_octree_pipeline_and_trace():
    """generate_voxel.py:75-150 end to end (dilation, bbox normalisation, strict filter, level), then the near/far
    tracer on the CUDA-built octree must equal the tracer restatement on the restated octree."""
    nrw = nrw_pkg()
    rng = np.random.RandomState(3)
    R = np.linalg.qr(rng.randn(3, 3))[0]
    sfm2gt = np.eye(4)
    sfm2gt[:3, :3] = R * 1.7
    sfm2gt[:3, 3] = [0.3, -0.2, 0.5]
    cfg = {"sfm2gt": sfm2gt.tolist(), "eval_bbx": [[-1.1, -0.9, -1.0], [1.2, 1.0, 0.8]]}
    gt_pts = op.sphere_shell_points(0.7, 0.05, n=3000, seed=5)
    gt2sfm = np.linalg.inv(sfm2gt)
    sfm_pts = gt_pts @ gt2sfm[:3, :3].T + gt2sfm[:3, 3]
    voxel = 0.05
    ref_tree, ref_origin, ref_scale, ref_level, _ = op.gen_octree(cfg, sfm_pts, voxel, expand=1)
    data = nrw.octree.make_octree_data(cfg, sfm_pts, voxel, device=0, expand=1)
    assert data["level"] == ref_level and data["scale"] == ref_scale
    assert np.array_equal(data["scene_origin"].cpu().numpy(), ref_origin)
    tree = {"octree": data["octree"], "prefix": data["spc_data"]["prefix"], "pyramid": data["spc_data"]["pyramid"],
            "points": data["spc_data"]["points"]}
    _cmp(tree, ref_tree)
    # dense occupancy (generate_voxel.py:181-186) of the leaf level
    dense = nrw.octree.convert_to_dense(tree, ref_level).cpu().numpy()
    leaves = ref_tree["levels"][ref_level]
    ref_dense = np.zeros_like(dense)
    ref_dense[leaves[:, 0], leaves[:, 1], leaves[:, 2]] = 1
    assert np.array_equal(dense, ref_dense)
    # trace through the C ABI with the built tensors
    from nrw import _lib
    L = _lib.lib()
    n = 256
    o = (rng.randn(n, 3) * 0.3 + ref_origin).astype(np.float32)
    d = rng.randn(n, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ro, rd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    near = torch.empty(n, device="cuda"); far = torch.empty(n, device="cuda")
    pid = torch.empty(n, dtype=torch.int32, device="cuda"); cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    pyr = tree["pyramid"].contiguous()
    so = (C.c_float * 3)(*[float(np.float32(v)) for v in ref_origin])
    _lib.check(L.nrw_octree_near_far(_lib.ptr(tree["octree"]), _lib.ptr(tree["prefix"]), C.c_void_p(pyr.data_ptr()), ref_level,
                                     _lib.ptr(ro), _lib.ptr(rd), n, so, float(ref_scale), _lib.ptr(near), _lib.ptr(far),
                                     _lib.ptr(pid), _lib.ptr(cnt), _lib.stream_ptr()), "nrw_octree_near_far")
    rn, rf, rp, rc = op.get_near_far(ref_tree, ref_level, o, d, ref_origin.astype(np.float32), np.float32(ref_scale))
    assert np.array_equal(cnt.cpu().numpy(), rc) and np.array_equal(pid.cpu().numpy(), rp)
    assert np.array_equal(near.cpu().numpy(), rn) and np.array_equal(far.cpu().numpy(), rf)
    assert (rc > 0).sum() > 20


def test_octree_refresh_matches_restatement():
    """neuconw_system.py:186-312 (surface_selection + octree_update) with an injected analytic SDF so that both sides
    threshold identical float32 values: candidate generation, dtypes, filtering and the rebuilt octree are bit-exact."""
    nrw = nrw_pkg()
    from util_nrw import build_system, synth
    cfg_scene = {"sfm2gt": np.eye(4).tolist(), "eval_bbx": [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]}
    pts = op.sphere_shell_points(0.55, 0.03, n=4000, seed=11)
    voxel = 0.12
    ref_tree, ref_origin, ref_scale, ref_level, _ = op.gen_octree(cfg_scene, pts, voxel, expand=1)
    pc = synth.PathConfig(n_samples=8, n_importance=8, up_sample_steps=1, n_outside=0)
    r = build_system(synth.make_params(seed=0), pc)["renderer"]
    r.scene_config, r.sfm_points, r.voxel_size = cfg_scene, pts, voxel
    r.octree_data = r.get_octree(0)
    assert r.octree_data["level"] == ref_level
    origin_sfm, radius_sfm = np.asarray(r.origin.cpu() if torch.is_tensor(r.origin) else r.origin, np.float32), r.radius

    def sdf_np(x):      # float32 sphere SDF, same op order on both sides
        return (np.sqrt((x.astype(np.float32) ** 2).sum(-1, dtype=np.float32)) - np.float32(0.5)).astype(np.float32)

    def sdf_t(x):
        return torch.from_numpy(sdf_np(x.cpu().numpy())).to(x.device)

    train_level, thr = ref_level + 2, 0.02
    new_ref, o_ref, s_ref, l_ref, tvs_ref, pc_ref = op.octree_update(cfg_scene, ref_tree, ref_origin, ref_scale, ref_level,
                                                                     train_level, thr, sdf_np, origin_sfm, radius_sfm)
    data = nrw.octree.octree_update(r, train_level, thr, sdf_fn=sdf_t)
    assert data["level"] == l_ref and data["voxel_size"] == tvs_ref and data["scale"] == s_ref
    assert r.fine_octree_data is data and len(pc_ref) > 100
    tree = {"octree": data["octree"], "prefix": data["spc_data"]["prefix"], "pyramid": data["spc_data"]["pyramid"],
            "points": data["spc_data"]["points"]}
    _cmp(tree, new_ref)
    # and with the real SDF network (nrw_sdf_query): the selection agrees with the restated fp32 network except for
    # candidates whose SDF is within 1e-4 of the threshold
    from util_nrw import port
    P = synth.make_params(seed=0)
    seen = {}

    def sdf_port(x):
        with torch.no_grad():
            v = port.sdf_forward(P, torch.from_numpy(np.ascontiguousarray(x)))[:, 0].numpy()
        seen["sdf"] = v
        return v

    op.surface_selection(ref_tree, ref_origin, ref_scale, ref_level, ref_level + 1, np.inf, sdf_port, origin_sfm, radius_sfm)
    thr2 = float(np.median(seen["sdf"]))
    pc_ref2, _ = op.surface_selection(ref_tree, ref_origin, ref_scale, ref_level, ref_level + 1, thr2, sdf_port, origin_sfm, radius_sfm)
    pc_gpu2, _ = nrw.octree.surface_selection(r, ref_level + 1, thr2)
    near_thr = int((np.abs(seen["sdf"] - thr2) < 1e-4).sum())
    assert len(pc_ref2) > 100 and abs(pc_gpu2.shape[0] - len(pc_ref2)) <= near_thr
    data2 = nrw.octree.octree_update(r, ref_level + 1, thr2)
    assert data2["octree"].numel() > 0 and data2["level"] >= ref_level
