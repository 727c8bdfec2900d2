"""CPU: self-consistency of the octree restatement (oracle/octree_port.py).  Kaolin itself is absent (parity UNPINNED),
so these pin the restatement to the SPC invariants its documentation states and that the reference's call sites rely on
(generate_voxel.py:173-186, 311-439)."""
import numpy as np

from oracle import octree_port as op


def test_octree_invariants():
    level = 6
    pts = op.sphere_shell_points(0.6, 0.04, n=5000, seed=1)
    t = op.build_octree(pts, level)
    pyr, octree, prefix, points = t["pyramid"], t["octree"], t["prefix"], t["points"]
    assert pyr.shape == (2, level + 2) and pyr[0, 0] == 1 and pyr[0, level + 1] == 0
    assert np.array_equal(pyr[1], np.concatenate([[0], np.cumsum(pyr[0])[:-1]]))
    n_nonleaf, n_total = int(pyr[1, level]), int(pyr[1, level + 1])
    assert len(octree) == n_nonleaf and len(points) == n_total
    pop = np.array([bin(int(b)).count("1") for b in octree])
    assert np.array_equal(prefix, np.concatenate([[0], np.cumsum(pop)[:-1]]))
    # children of hierarchy node i are nodes 1 + prefix[i] .. 1 + prefix[i] + popcount - 1, in Morton-digit order
    for i in range(n_nonleaf):
        kids = points[1 + prefix[i]:1 + prefix[i] + pop[i]]
        digits = [j for j in range(8) if octree[i] >> j & 1]
        want = np.array([[2 * points[i][0] + (j >> 2 & 1), 2 * points[i][1] + (j >> 1 & 1), 2 * points[i][2] + (j & 1)] for j in digits])
        assert np.array_equal(kids, want)
    # the leaf level is exactly the set of quantised input points
    q = np.unique(op.quantize_points(pts, level), axis=0)
    leaves = points[pyr[1, level]:pyr[1, level + 1]]
    assert set(map(tuple, q)) == set(map(tuple, leaves)) and len(q) == len(leaves)
    assert leaves.min() >= 0 and leaves.max() < 2 ** level


def test_gen_octree_and_refresh_shapes():
    cfg = {"sfm2gt": np.eye(4).tolist(), "eval_bbx": [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]}
    pts = op.sphere_shell_points(0.5, 0.03, n=800, seed=2)
    tree, origin, scale, level, pf = op.gen_octree(cfg, pts, 0.1, expand=1)
    assert level == int(np.floor(np.log2(2 * scale / 0.1))) and np.allclose(origin, 0) and scale == 1.0
    assert (np.abs(pf) < 1).all() and len(pf) > len(pts)            # 3x3x3 dilation adds points, strict cube filter
    sdf = lambda x: np.linalg.norm(x, axis=-1).astype(np.float32) - np.float32(0.5)
    new_tree, o2, s2, l2, tvs, pc = op.octree_update(cfg, tree, origin, scale, level, level + 1, 0.03, sdf,
                                                     np.zeros(3, np.float32), 1.0)
    assert tvs == 2 / 2 ** (level + 1) * scale and l2 in (level, level + 1)
    assert len(pc) > 0 and float((np.linalg.norm(pc, axis=-1) - 0.5).max()) <= 0.03 + 1e-6   # sdf <= threshold kept
    # a ray through the centre hits the refreshed octree on both sides of the sphere
    near, far, pid, cnt = op.get_near_far(new_tree, l2, np.array([[-2.0, 0.01, 0.02]], np.float32), np.array([[1.0, 0, 0]], np.float32),
                                          o2.astype(np.float32), np.float32(s2))
    assert cnt[0] >= 2 and 1.3 < near[0] < 1.7 and 2.3 < far[0] < 2.7


def test_octree_to_spc_decoder_matches_restatement():
    """nrw.generate_voxel.octree_to_spc (torch decode of the child-mask bytes; stand-in for Kaolin's scan_octrees +
    generate_points with the reference's signature, generate_voxel.py:173-178) vs the restatement's tables."""
    import torch
    from nrw.generate_voxel import convert_to_dense, octree_to_spc

    for level, seed in ((5, 2), (7, 3)):
        pts = op.sphere_shell_points(0.55, 0.05, n=3000, seed=seed)
        t = op.build_octree(pts, level)
        points, pyramid, prefix = octree_to_spc(torch.from_numpy(t["octree"].astype(np.uint8)))
        assert np.array_equal(points.numpy(), t["points"].astype(np.int16))
        assert np.array_equal(pyramid.numpy(), t["pyramid"].astype(np.int32))
        assert np.array_equal(prefix.numpy(), t["prefix"].astype(np.int32))
        dense = convert_to_dense(torch.from_numpy(t["octree"].astype(np.uint8)), level)
        leaves = t["points"][t["pyramid"][1, level]:t["pyramid"][1, level + 1]]
        assert dense.shape == (2 ** level,) * 3 and int(dense.sum()) == len(leaves)
        assert bool((dense[leaves[:, 0], leaves[:, 1], leaves[:, 2]] == 1).all())
