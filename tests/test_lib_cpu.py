"""CPU-side checks of the C-ABI library and the host mirror (no GPU, no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import synth


def test_library_loads_and_exports_every_declared_symbol():
    from nrw import _lib

    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "nrw.h")).read()
    declared = set(re.findall(r"NRW_API\s+[\w\s\*]+?\b(nrw_\w+)\s*\(", hdr))
    assert declared, "no NRW_API declarations found"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nrw_version() >= 100


def test_param_table_matches_reference_checkpoint_layout():
    """names / shapes of SURVEY.md 9.4 (pinned to the reference by tests/test_oracle_vs_reference.py)."""
    from nrw import _lib

    tab, total = _lib.param_table(5000, 48)
    P = synth.make_params(seed=0)
    assert {n: tuple(s) for n, s, _, _ in tab} == {k: tuple(v.shape) for k, v in P.items()}
    offs = sorted((o, n) for _, _, o, n in tab)
    for (o0, n0), (o1, _) in zip(offs, offs[1:]):
        assert o0 + n0 <= o1 and o1 % 4 == 0
    assert total >= offs[-1][0] + offs[-1][1]
    assert sum(n for _, _, _, n in tab) == 3896255 + 0  # embedding 240000 + neuconw 2957627 + nerf 698628


def test_ctypes_struct_layouts():
    from nrw import _lib

    assert ctypes.sizeof(_lib.SamplerCfg) == 28
    assert ctypes.sizeof(_lib.RenderCfg) == 32
    assert ctypes.sizeof(_lib.RenderIO) == 8 * len(_lib._IO_FIELDS)
    assert ctypes.sizeof(_lib.RenderGrads) == 8 * len(_lib._GRAD_FIELDS)
    assert ctypes.sizeof(_lib.ParamInfo) == 32


def test_modules_are_checkpoint_compatible():
    import nrw
    from util_nrw import COLOR_CONFIG, SDF_CONFIG

    P = synth.make_params(seed=0)
    m = nrw.NeuconW(SDF_CONFIG, COLOR_CONFIG, dict(init_val=0.3), in_channels_a=48, encode_a=True)
    n = nrw.NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4],
                 encode_appearance=True, in_channels_a=48, in_channels_dir=27, use_viewdirs=True)
    m.load_state_dict({k[len("neuconw."):]: v for k, v in P.items() if k.startswith("neuconw.")}, strict=True)
    n.load_state_dict({k[len("nerf."):]: v for k, v in P.items() if k.startswith("nerf.")}, strict=True)
    # geometric init statistics (models/neuconw.py:222-254)
    fresh = nrw.NeuconW(SDF_CONFIG, COLOR_CONFIG, dict(init_val=0.3), in_channels_a=48, encode_a=True)
    w8 = fresh.sdf_net.lin8.weight_v
    assert abs(float(w8.mean()) - (3.14159265 ** 0.5) / (512 ** 0.5)) < 1e-3
    assert float(fresh.sdf_net.lin0.weight_v[:, 3:].abs().max()) == 0.0
    assert float(fresh.sdf_net.lin8.bias.mean()) == -0.5


def test_unsupported_configurations_fail_loudly():
    import nrw
    from util_nrw import COLOR_CONFIG, SDF_CONFIG

    bad = dict(SDF_CONFIG, d_hidden=256)
    with pytest.raises(nrw.NrwError):
        nrw.NeuconW(bad, COLOR_CONFIG, dict(init_val=0.3), in_channels_a=48, encode_a=True)
    with pytest.raises(nrw.NrwError):
        nrw.NeRF(D=4)


def test_render_refuses_cpu_tensors():
    import nrw
    from util_nrw import build_system

    cfg = synth.PathConfig(n_samples=8, n_importance=8, up_sample_steps=2)
    s = build_system(synth.make_params(0), cfg, device="cpu")
    b = synth.make_rays(4, cfg)
    with pytest.raises(nrw.NrwError):
        s["renderer"].render(b["rays"], b["ts"], b["label"])


def test_octree_and_optimizer_entry_points_fail_loudly_without_cuda():
    """K0 / refresh / fused optimiser have no CPU path either: clear errors, no silent fallback."""
    import nrw
    import nrw.octree as noct
    from util_nrw import build_system

    with pytest.raises(nrw.NrwError):
        noct.build_octree(torch.zeros(8, 3, dtype=torch.float64), 4)          # CPU tensor
    cfg = synth.PathConfig(n_samples=8, n_importance=8, up_sample_steps=2)
    r = build_system(synth.make_params(0), cfg, device="cpu")["renderer"]
    with pytest.raises(nrw.NrwError):
        r.get_octree("cpu")                                                   # no sfm_points / scene_config given
    with pytest.raises(nrw.NrwError):
        noct.octree_update(r, 5, 0.01)                                        # no scene_config
    # argument validation of the C entry points happens before any CUDA call
    from nrw import _lib
    L = _lib.lib()
    assert L.nrw_octree_build(None, 0, 4, 20, None, None, None, None, 1, 1, None, None, None) != 0
    assert b"octree_build" in L.nrw_last_error()
    assert L.nrw_adam_clip_step(None, None, None, None, 5, None, 0.99, 1e-4, 0.9, 0.999, 1e-7, 1, None) != 0
    assert L.nrw_boundary_samples(4, 0, 2, None, None, None, None, None) != 0
    assert L.nrw_octree_build_scratch_bytes(1000, 6, 6000) > 0
