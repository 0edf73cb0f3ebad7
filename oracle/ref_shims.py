"""Oracle tooling (test infrastructure, not product): import the UNMODIFIED
reference sources from ``/root/reference`` in the build container.

Only ``oracle/make_golden.py`` and the optional ``reference``-marked CPU tests
use this, to pin the restatements in ``oracle/*_ref.py`` against the reference
itself.  ``/root/reference`` does not exist on the GPU box; nothing run there
may import this module.

The released reference has bit-rotted against this container (SURVEY.md
section 0, D6/D7): removed numpy aliases, ``torch._six``, and third-party
packages that are not installed (timm 0.4.12, matplotlib, tensorboardX, mmcv,
...).  The shims below restore exactly those names; no reference file is
edited or copied.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MEM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mem"))


class _Anything:
    """Attribute sink for off-path names the reference imports but the hot path never calls."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


_STUB_ROOTS = ("timm", "matplotlib", "tensorboardX", "mmcv", "configargparse", "h5py",
               "deepspeed", "dall_e", "apex", "horovod", "wandb")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False
_registry: dict = {}


def install() -> None:
    """Idempotently install the compatibility shims."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy as np
    import torch

    # numpy aliases removed in numpy >= 1.24 (datasets.py:568, masking_generator.py:69, ...)
    for alias, typ in (("int", int), ("float", float), ("bool", bool)):
        if alias not in np.__dict__:
            setattr(np, alias, typ)
    # torch._six was removed (utils.py:25)
    six = types.ModuleType("torch._six")
    six.inf = math.inf
    sys.modules.setdefault("torch._six", six)
    torch._six = sys.modules["torch._six"]

    # scipy.interpolate.interp2d was removed in SciPy 1.14 (utils.py:696 calls it with kind='cubic' on a regular grid);
    # SciPy's removal notice names the replacement used here: RectBivariateSpline on the same grid, transposed.
    try:
        from scipy import interpolate as _si

        def interp2d(x, y, z, kind="linear"):
            k = {"linear": 1, "cubic": 3, "quintic": 5}[kind]
            spline = _si.RectBivariateSpline(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64),
                                             np.asarray(z, dtype=np.float64).T, kx=k, ky=k, s=0)
            return lambda xn, yn: spline(np.asarray(xn, dtype=np.float64), np.asarray(yn, dtype=np.float64)).T
        try:
            _si.interp2d([0, 1, 2, 3], [0, 1, 2, 3], np.zeros((4, 4)), kind="cubic")
        except NotImplementedError:
            _si.interp2d = interp2d
    except ImportError:
        pass

    # Only stub packages that are genuinely missing.
    missing = []
    for root in _STUB_ROOTS:
        try:
            importlib.import_module(root)
        except Exception:
            missing.append(root)
    globals()["_STUB_ROOTS"] = tuple(missing)
    sys.meta_path.append(_StubFinder())

    if "timm" in missing:
        # timm 0.4.12 pieces that ARE on the hot path: restated here.
        layers = importlib.import_module("timm.models.layers")

        def drop_path(x, drop_prob: float = 0.0, training: bool = False):
            if drop_prob == 0.0 or not training:
                return x
            keep = 1 - drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            gate = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
            gate.floor_()
            return x.div(keep) * gate

        def to_2tuple(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        layers.drop_path, layers.to_2tuple, layers.trunc_normal_ = drop_path, to_2tuple, trunc_normal_
        registry = importlib.import_module("timm.models.registry")

        def register_model(fn):
            _registry[fn.__name__] = fn
            return fn

        registry.register_model = register_model
        models = importlib.import_module("timm.models")
        models.create_model = lambda name, **kw: _registry[name](**kw)
        tutils = importlib.import_module("timm.utils")
        tutils.get_state_dict = lambda m, *a, **k: m.state_dict()
    if "tensorboardX" in missing:
        importlib.import_module("tensorboardX").SummaryWriter = object

    for sub in ("mem", "eventvae", ""):
        p = os.path.join(REFERENCE_ROOT, sub) if sub else REFERENCE_ROOT
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def ref_module(name: str):
    """Import a reference module by its in-tree name, e.g. ``masking_generator``,
    ``modeling_pretrain``, ``vae.vae_model``, ``datasets``, ``engine_for_pretraining``."""
    install()
    return importlib.import_module(name)


def ref_create_model(name: str, **kwargs):
    install()
    ref_module("modeling_pretrain")
    ref_module("modeling_finetune")
    return _registry[name](**kwargs)
