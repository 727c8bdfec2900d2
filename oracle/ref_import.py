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
                """the slice of pytorch_lightning.LightningModule NeuconWSystem touches (neuconw_system.py)"""

                def __init__(self, *a, **k):
                    super().__init__()
                    self.global_step = 0
                    self.logged = {}
                    self.trainer = types.SimpleNamespace(global_rank=0, save_checkpoint=lambda *a, **k: None)
                    self.logger = types.SimpleNamespace(save_dir="/tmp", name="nrw", experiment=mock.MagicMock())

                def save_hyperparameters(self, hparams=None, *a, **k):
                    self.hparams = hparams

                def log(self, name, value, *a, **k):
                    self.logged[name] = value

                @property
                def device(self):
                    return next(self.parameters()).device

            pl.LightningModule = LightningModule
            pl.LightningDataModule = object
            pl.seed_everything = lambda s: torch.manual_seed(s)
            sys.modules["pytorch_lightning"] = pl


class CfgNode(dict):
    """Minimal yacs.config.CfgNode (attribute access, clone, merge_from_file with literal evaluation of strings such
    as "(4,)") for config/defaults.py when yacs is not installed."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def _merge(self, other):
        import ast
        for k, v in other.items():
            if isinstance(v, dict):
                node = self.get(k)
                if not isinstance(node, CfgNode):
                    node = self[k] = CfgNode()
                node._merge(v)
            else:
                if isinstance(v, str):
                    try:
                        v = ast.literal_eval(v)
                    except (ValueError, SyntaxError):
                        pass
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f))


def _install_yacs():
    try:
        import yacs.config  # noqa: F401
        if not isinstance(sys.modules.get("yacs"), mock.MagicMock):
            return
    except Exception:
        pass
    y, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")
    yc.CfgNode = CfgNode
    y.config = yc
    sys.modules["yacs"], sys.modules["yacs.config"] = y, yc


def load_system():
    """The reference's LightningModule (lightning_modules/neuconw_system.py) and config defaults, importable without
    pytorch_lightning / yacs / kaolin: returns a namespace with the module `ns` (so callers can re-bind
    ns.NeuconW / ns.NeRF / ns.NeuconWRenderer / ns.gen_octree ... - the documented drop-in patch), NeuconWSystem,
    get_cfg_defaults and load_ckpt (utils/__init__.py:64-98)."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    _install_yacs()
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lightning_modules.neuconw_system as ns  # type: ignore
        from config.defaults import get_cfg_defaults  # type: ignore
        from utils import load_ckpt  # type: ignore
    return types.SimpleNamespace(ns=ns, NeuconWSystem=ns.NeuconWSystem, get_cfg_defaults=get_cfg_defaults, load_ckpt=load_ckpt,
                                 config_dir=os.path.join(REF_ROOT, "config"))


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
