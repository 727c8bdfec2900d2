"""GPU: data movers either side of the hot path (csrc/dataio.cu through the C ABI) - BIT-EXACT against the
restatements in oracle/dataio_port.py (which tests/test_dataio_oracle.py pins to the unmodified reference):
ray-cache gather + RAY_MASK_LIST filter, RayCache epoch semantics and prefetch, dense / sparse query lattices,
threshold compaction, the SDF-volume pipeline of mesh extraction (BASELINE config 5 path)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import dataio_port as dp
from util_nrw import build_system, synth

pytestmark = pytest.mark.gpu


def _cache(n, seed=0, masked=True):
    g = torch.Generator().manual_seed(seed)
    rays = torch.randn(n, 12, generator=g)
    rays[:, 8] = torch.randint(0, 1500, (n,), generator=g).float()
    labels = [0.0, 2.0, 6.0, 1.0] + ([12.0, 20.0, 116.0, 127.0] if masked else [])
    rays[:, 9] = torch.tensor(labels)[torch.randint(0, len(labels), (n,), generator=g)]
    return rays.contiguous(), torch.rand(n, 3, generator=g)


@pytest.mark.parametrize("n,B,masked", [(5000, 1024, True), (5000, 777, True), (300, 300, False), (4096, 1, True), (70000, 65536, True)])
def test_raycache_gather_bitexact(n, B, masked):
    from nrw.raycache import RayCache

    rays, rgbs = _cache(n, seed=n + B, masked=masked)
    rc = RayCache(rays, rgbs, batch_size=B, device="cuda", prefetch=False)
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, n, (B,), generator=g)
    buf = rc.gather(idx.cuda())
    torch.cuda.synchronize()
    m = int(buf["n_valid"])
    want = dp.filter_batch(dp.getitem_batch(rays, rgbs, idx))
    assert m == want["rays"].shape[0]
    assert torch.equal(buf["rays"][:m].cpu(), want["rays"])
    assert torch.equal(buf["rgbs"][:m].cpu(), want["rgbs"])
    assert torch.equal(buf["ts"][:m].cpu(), want["ts"])
    assert torch.equal(buf["label"][:m].cpu(), want["label"])


def test_raycache_all_rows_masked_and_empty():
    from nrw.raycache import RayCache

    rays, rgbs = _cache(64, masked=False)
    rays[:, 9] = 12.0                              # every ray is a person
    rc = RayCache(rays, rgbs, batch_size=32, device="cuda", prefetch=False)
    buf = rc.gather(torch.arange(32, device="cuda"))
    assert int(buf["n_valid"]) == 0
    b = rc.next_batch()
    assert b["n_valid"] == 0 and b["rays"].shape == (0, 10)


def test_raycache_epoch_semantics_and_prefetch():
    """RandomSampler + DataLoader(batch_size, drop_last=False): every row exactly once per epoch, short last batch."""
    from nrw.raycache import RayCache

    n, B = 1000, 256
    rays, rgbs = _cache(n, masked=False)
    rays[:, 0] = torch.arange(n).float()           # row id travels in column 0
    rc = RayCache(rays, rgbs, batch_size=B, device="cuda", ray_mask_list=None, seed=3, prefetch=True)
    assert len(rc) == 4
    for epoch in range(2):
        seen = []
        sizes = []
        for _ in range(len(rc)):
            b = rc.next_batch()
            sizes.append(b["n_valid"])
            seen.append(b["rays"][:, 0].clone())
            assert torch.equal(b["rgbs"].cpu(), rgbs[b["rays"][:, 0].long().cpu()])
        assert sizes == [256, 256, 256, 232]
        ids = torch.cat(seen).cpu().long()
        assert torch.equal(torch.sort(ids)[0], torch.arange(n))
    # prefetch on / off deliver the same stream of batches for the same seed
    a = RayCache(rays, rgbs, batch_size=B, device="cuda", seed=9, prefetch=True)
    c = RayCache(rays, rgbs, batch_size=B, device="cuda", seed=9, prefetch=False)
    for _ in range(6):
        ba, bc = a.next_batch(), c.next_batch()
        assert ba["n_valid"] == bc["n_valid"] and torch.equal(ba["rays"], bc["rays"]) and torch.equal(ba["ts"], bc["ts"])


def test_synthetic_cache_feeds_training_step():
    from nrw.raycache import RayCache, synthetic_cache
    from nrw.train import TrainSystem

    rays, rgbs = synthetic_cache(20000, n_images=16, n_vocab=64, seed=2, masked_fraction=0.1)
    rc = RayCache(rays, rgbs, batch_size=512, device="cuda", seed=1)
    sysm = TrainSystem(torch.device("cuda", 0), n_samples=16, n_importance=8, up_sample_steps=2, n_outside=4, n_vocab=64,
                       chunk_rows=8192, batch_size=512)
    for _ in range(3):
        b = rc.next_batch()
        assert 380 < b["n_valid"] < 512                          # ~10 % of the rows are person / car
        assert not bool(((b["label"] == 12) | (b["label"] == 20)).any())
        loss = sysm.training_step(b)
        assert torch.isfinite(loss)


# ------------------------------------------------------------------------------------------------------------
def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


@pytest.mark.parametrize("dim,origin,radius", [(17, (0.0, 0.0, 0.0), 1.0), (33, (0.3, -0.2, 0.1), 0.37), (2, (0.0, 0.0, 0.0), 1.0)])
def test_dense_lattice_bitexact(dim, origin, radius):
    from nrw import _lib

    L = _lib.lib()
    want = dp.dense_lattice(dim, origin, radius)
    n = dim ** 3
    out = torch.zeros(n, 3, device="cuda")
    o64 = np.array(origin, dtype=np.float64)
    lo, hi = (o64 - radius).astype(np.float32), (o64 + radius).astype(np.float32)
    half = n // 2                                               # two chunks with an odd split
    _lib.check(L.nrw_grid_points_dense(dim, _f3(lo), _f3(hi), 0, half, _lib.ptr(out), _lib.stream_ptr()), "dense")
    _lib.check(L.nrw_grid_points_dense(dim, _f3(lo), _f3(hi), half, n - half, C.c_void_p(out.data_ptr() + half * 12),
                                       _lib.stream_ptr()), "dense")
    torch.cuda.synchronize()
    # torch's CPU linspace evaluates VECTOR lanes as (start + step*i0) + step*j (two roundings), the kernel evaluates every
    # element with the scalar formula of ATen (one multiply, one add): equal up to 1 ulp of the coordinate, exact for the
    # lattice index -> coordinate mapping (order, count)
    got = out.cpu()
    ulp = torch.maximum(want.abs(), torch.full_like(want, float(radius))) * 2.0 ** -23
    assert bool(((got - want).abs() <= ulp).all())
    assert torch.equal(got[:, 0].reshape(dim, dim, dim)[:, 0, 0], got[::dim * dim, 0])          # x slowest, z fastest
    sym = torch.linspace(float(lo[2]), float(hi[2]), dim)
    assert float((got[:dim, 2] - sym).abs().max()) <= float(ulp.max())


def test_sparse_lattice_and_threshold_compaction_bitexact():
    from nrw import _lib

    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    m, up = 213, 4
    ind = torch.unique(torch.randint(0, 32, (m, 3), generator=g), dim=0)       # lexicographically sorted = nonzero order
    voxel = 2 / (2 ** 7) * 3.7
    vol_origin = torch.tensor([0.11, -0.52, 6.3]) - 3.7
    scene_origin = torch.tensor([0.568699, -0.0935532, 6.28958])
    radius = 4.6
    xs_want, xt_want = dp.sparse_lattice(ind, up, voxel, vol_origin, scene_origin, radius)
    n = ind.shape[0] * up ** 3
    leaves = ind.to(torch.int16).cuda().contiguous()
    xs, xt = torch.zeros(n, 3, device="cuda"), torch.zeros(n, 3, device="cuda")
    _lib.check(L.nrw_grid_points_sparse(_lib.ptr(leaves), ind.shape[0], up, float(np.float32(voxel)), _f3(vol_origin), _f3(scene_origin),
                                        radius, 0, n, _lib.ptr(xs), _lib.ptr(xt), _lib.stream_ptr()), "sparse")
    torch.cuda.synchronize()
    assert torch.equal(xs.cpu().view(torch.int32), xs_want.view(torch.int32))
    assert torch.equal(xt.cpu().view(torch.int32), xt_want.view(torch.int32))
    # xyz_sfm[sdf <= threshold], appended over three ragged chunks
    sdf = torch.randn(n, generator=g).cuda()
    thr = 0.1
    out = torch.zeros(n, 3, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    sb = L.nrw_compact_scratch_bytes(n)
    scratch = torch.zeros(sb + 256, dtype=torch.uint8, device="cuda")
    sp = (scratch.data_ptr() + 255) // 256 * 256
    cuts = [0, 1000, 1001, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        _lib.check(L.nrw_threshold_compact(C.c_void_p(sdf.data_ptr() + a * 4), C.c_void_p(xs.data_ptr() + a * 12), b - a, thr,
                                           _lib.ptr(out), _lib.ptr(count), C.c_void_p(sp), _lib.stream_ptr()), "compact")
    torch.cuda.synchronize()
    want = xs_want[(sdf.cpu() <= thr)]
    assert int(count) == want.shape[0]
    assert torch.equal(out[:int(count)].cpu(), want)


def test_sdf_volume_matches_pointwise_queries():
    """utils/visualization.py:42-96 dense branch: the device pipeline returns what renderer.sdf gives on the torch lattice."""
    from nrw.mesh import sdf_volume

    P = synth.make_params(seed=0)
    cfg = synth.PathConfig()
    r = build_system(P, cfg, precision="bf16x3", backend=0, chunk_rows=8192)["renderer"]
    dim = 24
    vol, vol_origin, voxel = sdf_volume(r, dim, origin=(0.1, 0.0, -0.05), radius=0.9, chunk=5000)
    # exact against pointwise queries on the kernel's own lattice (chunking invariance of the pipeline) ...
    from nrw import _lib
    L = _lib.lib()
    o64 = np.array((0.1, 0.0, -0.05), dtype=np.float64)
    lo, hi = (o64 - 0.9).astype(np.float32), (o64 + 0.9).astype(np.float32)
    pts = torch.zeros(dim ** 3, 3, device="cuda")
    _lib.check(L.nrw_grid_points_dense(dim, _f3(lo), _f3(hi), 0, dim ** 3, _lib.ptr(pts), _lib.stream_ptr()), "dense")
    own = r.sdf(pts.reshape(-1, 1, 3)).reshape(-1)
    assert vol.shape == (dim, dim, dim) and torch.equal(vol.reshape(-1), own), float((vol.reshape(-1) - own).abs().max())
    # ... and equal to the queries on torch's CPU-built lattice up to the effect of its 1-ulp coordinate differences
    # (see test_dense_lattice_bitexact; sin(32 x) amplifies an ulp of x by 32)
    want = r.sdf(dp.dense_lattice(dim, (0.1, 0.0, -0.05), 0.9).cuda().reshape(-1, 1, 3)).reshape(dim, dim, dim)
    assert float((vol - want).abs().max()) < 5e-4, float((vol - want).abs().max())
    assert abs(voxel - 2 * 0.9 / (dim - 1)) < 1e-12 and np.allclose(vol_origin, np.array([0.1, 0.0, -0.05]) - 0.9)
