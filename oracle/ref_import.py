"""Import the UNMODIFIED reference modules from /root/reference (build container) or from the verbatim copy
``oracle/_ref`` made by ``oracle/fetch_ref.py`` (git-ignored; travels to the GPU box).

TEST INFRASTRUCTURE.  The reference imports ~10 third-party packages at module
scope that are absent here (open3d, kaolin, pytorch_lightning, ...).  None of
them is touched on the cached-near/far training path, so they
are replaced by inert stand-ins in ``sys.modules`` before the import.

``/root/reference`` does not exist on the GPU box; there ``oracle/_ref`` (sha256-verified against its
manifest) is used.  Callers must check ``available()`` first.
"""
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root():
    env = os.environ.get("NRW_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/rendering/renderer.py"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF_ROOT = _resolve_root()

_STUBS = [
    "open3d", "kaolin", "kaolin.ops", "kaolin.ops.spc", "kaolin.render",
    "kaolin.render.spc", "torch_optimizer", "trimesh", "skimage",
    "skimage.measure", "kornia", "kornia.losses", "h5py", "lpips", "matplotlib",
    "matplotlib.pyplot", "matplotlib.colors", "yacs", "yacs.config", "loguru",
    "mcubes", "cv2", "PIL", "PIL.Image", "torchvision", "torchvision.transforms",
    "imageio", "plyfile", "ray", "pyrender", "test_tube",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "rendering", "renderer.py"))


def _install_stubs():
    import torch

    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        sys.modules[name] = mock.MagicMock(name=name)
    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:
            pl = types.ModuleType("pytorch_lightning")

            class LightningModule(torch.nn.Module):
                def save_hyperparameters(self, *a, **k):
                    pass

                def log(self, *a, **k):
                    pass

            pl.LightningModule = LightningModule
            pl.LightningDataModule = object
            pl.seed_everything = lambda s: torch.manual_seed(s)
            sys.modules["pytorch_lightning"] = pl


def load():
    """Returns a namespace with the reference's NeuconW, NeRF, NeuconWRenderer, NeuconWLoss."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.neuconw import NeuconW  # type: ignore
        from models.nerf import NeRF  # type: ignore
        from rendering.renderer import NeuconWRenderer  # type: ignore
        from losses import NeuconWLoss  # type: ignore
    ns = types.SimpleNamespace(NeuconW=NeuconW, NeRF=NeRF, NeuconWRenderer=NeuconWRenderer,
                               NeuconWLoss=NeuconWLoss)
    return ns
