"""BASELINE config 5 (sdf_extract / marching-cubes SDF grid): the batched SDF query (nrw_sdf_query, renderer.sdf,
renderer.py:947-949) over a dense grid.  Parity on a sample of the grid against the restated network, plus the
size-independent properties: the result does not depend on how the grid is split into batches or engine chunks."""
import numpy as np
import pytest
import torch

from util_nrw import build_system, port, synth

pytestmark = pytest.mark.gpu


def test_sdf_grid_query():
    P = synth.make_params(seed=0)
    r = build_system(P, synth.PathConfig(), precision="bf16x3", backend=0, chunk_rows=65536)["renderer"]
    n = 96                                                      # 96^3 = 884,736 points (level-10 extraction uses 512^3)
    ax = torch.linspace(-1.0, 1.0, n, device="cuda")
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 1, 3)
    with torch.no_grad():
        full = r.sdf(grid)                                      # one call, engine chunks of 65536 rows
        parts = torch.cat([r.sdf(grid[i:i + 100003]) for i in range(0, grid.shape[0], 100003)])   # ragged batches
    assert full.shape == (n ** 3, 1) and torch.isfinite(full).all()
    assert torch.equal(full, parts)                             # batching / chunking invariance (row-independent GEMMs)
    idx = torch.randint(0, n ** 3, (4096,), generator=torch.Generator().manual_seed(1))
    ref = port.sdf_value(P, grid[idx.cuda()].reshape(-1, 3).cpu()).detach()
    got = full[idx.cuda()].cpu()
    assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert float(full.std()) > 0.0
