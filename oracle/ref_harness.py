"""Import harness around the UNMODIFIED reference tree (test infrastructure only).

This file is part of the oracle: it may be imported only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg.  It only works in a
container that has /root/reference mounted (the GPU box has not) and is used to
(a) pin oracle/dpn_oracle.py against the reference's own code and (b) generate the
golden vectors committed under tests/golden/ (see oracle/make_golden.py).

The reference imports GDAL / xarray / wrf / ... at module import time
(utils/gdal_utils.py:5-15, utils/downscale_utils.py:10-15, dataset/physics_dataset.py:21,
utils/utils.py:7-13, train.py:4).  None of these touch the hot path, so they are replaced by
empty ModuleType stubs (with __spec__, or torch._dynamo chokes on them).
"""
import importlib.machinery
import os
import runpy
import sys
import types

import torch

REF_ROOT = os.environ.get("DPN_REFERENCE_ROOT", "/root/reference")

_STUBBED = ["osgeo", "osgeo.gdal", "osgeo.osr", "gdal", "osr", "pyproj", "bs4", "netCDF4", "wrf",
            "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "xarray", "skimage", "mmcv", "cv2",
            "mpl_toolkits", "mpl_toolkits.basemap", "tqdm"]


class _Stub(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return _Stub(self.__name__ + "." + key)

    def __call__(self, *a, **k):
        return None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "DeepPhysiNet"))


def _install_stubs():
    for name in _STUBBED:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = _Stub(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
        mod.__path__ = []
        sys.modules[name] = mod


def load_reference():
    """Returns (InterfacePhysics class, builder_loss, config dict)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from DeepPhysiNet.interface.interface_physics import InterfacePhysics
    from DeepPhysiNet.losses.builder import builder_loss
    cfg = dict(runpy.run_path(os.path.join(REF_ROOT, "configs", "DeepPhysiNet_NCEP_cfg.py"))["config"])
    cfg.pop("name")
    return InterfacePhysics, builder_loss, cfg


def build_reference_model(seed=0, dtype=torch.float32, dx=27000.0, dy=27000.0, img_size=(145, 257),
                          pred_t_span=86400, with_clip=True):
    """Constructs InterfacePhysics exactly as train.py:39 does and sets the attributes that
    run_train_interface sets before the loop (interface_physics.py:339-342, :418, :438)."""
    InterfacePhysics, builder_loss, cfg = load_reference()
    cfg["train_cfg"] = dict(cfg["train_cfg"])
    cfg["train_cfg"]["img_size"] = tuple(img_size)
    torch.manual_seed(seed)
    m = InterfacePhysics(**cfg)
    m.dx, m.dy, m.dt = float(dx), float(dy), 3600.0
    m.pred_t_span = pred_t_span
    m.with_clip = with_clip
    if dtype == torch.float64:
        m = m.double()
    return m, builder_loss, cfg


class fp64_mode:
    """The reference hard-casts `.float()` inside every loss (interface_physics.py:104,114,...);
    neutralise it in this process only so the reference can run in fp64."""

    def __enter__(self):
        self._orig = torch.Tensor.float
        torch.Tensor.float = lambda s, *a, **k: s
        return self

    def __exit__(self, *exc):
        torch.Tensor.float = self._orig
        return False
