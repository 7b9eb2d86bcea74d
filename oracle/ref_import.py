"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (read-only checkout at /root/reference) on CPU.

Only `oracle/gen_golden.py` (run in the build container, where /root/reference exists) uses this, to pin the
oracle restatement and to produce the committed fixtures under tests/golden/.  Nothing in the product, the GPU
tests, smoke() or bench.py imports it (the reference checkout does not exist on the GPU box).

The shims below are the ones SURVEY.md §8c lists; none of them changes reference arithmetic:
  * ftfy / torchtyping / tensorboardX / h5py / transforms3d / skimage / pybullet / matplotlib / ray / imageio:
    absent optional imports -> stub modules;
  * torch_scatter.scatter: absent third-party op (pytorch-scatter 2.0.9, semabs.yml:112) -> restated with
    torch.Tensor.scatter_reduce (mean / amax, empty cells = 0);
  * typeguard.typechecked -> no-op (installed typeguard 4 cannot evaluate torchtyping string annotations);
  * np.NAN alias (numpy 2 removed it);
  * CLIP/clip/__init__.py:227 indexes a tensor with an object ndarray of slices, which torch >= 2 rejects ->
    create_tiles is wrapped so each row iterates as a tuple.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_SNAPSHOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/build_ref.py's byte-for-byte copy


def _default_root():
    if os.environ.get("SEMABS_REFERENCE_ROOT"):
        return os.environ["SEMABS_REFERENCE_ROOT"]
    if os.path.isdir("/root/reference/CLIP/clip"):
        return "/root/reference"
    return _SNAPSHOT  # the GPU box: only the snapshot exists (it has the path's files only: CLIP/clip, net, unet3d, lamb)


REF_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "CLIP", "clip"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    if not attrs and "." not in name:
        # only stand in for what is really absent: a stub of an installed package (filelock, imageio ...) would break every
        # later import of it in the same process
        try:
            if importlib.util.find_spec(name) is not None:
                return importlib.import_module(name)
        except (ImportError, ValueError):
            pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter semantics for the one call site net.py:193-200 (dim=-2, index broadcast over the
    trailing feature axis)."""
    assert out is None
    dim = dim % src.dim()
    idx = index
    while idx.dim() < src.dim():
        idx = idx.unsqueeze(-1)
    idx = idx.expand_as(src)
    shape = list(src.shape)
    shape[dim] = int(dim_size) if dim_size is not None else int(index.max()) + 1
    red = {"mean": "mean", "max": "amax", "sum": "sum", "add": "sum", "min": "amin"}[reduce]
    return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_reduce(dim, idx, src, reduce=red, include_self=False)


def install_shims():
    _stub("ftfy", fix_text=lambda s: s)

    class _TT:
        def __class_getitem__(cls, item):
            return torch.Tensor

    _stub("torchtyping", TensorType=_TT, patch_typeguard=lambda: None)
    import typeguard

    typeguard.typechecked = lambda f=None, **kw: (f if f is not None else (lambda g: g))
    _stub("torch_scatter", scatter=_scatter)
    if not hasattr(np, "NAN"):
        np.NAN = np.nan
    for name in ("tensorboardX", "h5py", "transforms3d", "pybullet", "pybullet_data", "ray", "imageio", "filelock"):
        _stub(name)
    sys.modules["tensorboardX"].SummaryWriter = object
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.use = lambda *a, **k: None


_clip_pkg = None


def import_reference_clip():
    """Returns the reference's `CLIP.clip` package (ClipWrapper, ClipGradcam, ...)."""
    global _clip_pkg
    if _clip_pkg is not None:
        return _clip_pkg
    assert reference_available(), f"reference checkout not found at {REF_ROOT}"
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pkg = importlib.import_module("CLIP.clip")
    orig_create_tiles = pkg.ClipWrapper.create_tiles.__func__

    class _TupleRows(np.ndarray):
        def __iter__(self):
            for i in range(self.shape[0]):
                row = np.ndarray.__getitem__(self, i)
                yield tuple(np.asarray(row).tolist()) if self.ndim > 1 else row

    def create_tiles(cls, *a, **k):
        tiles, tile_imgs, counts, tile_sizes = orig_create_tiles(cls, *a, **k)
        return tiles.view(_TupleRows), tile_imgs, counts, tile_sizes

    pkg.ClipWrapper.create_tiles = classmethod(create_tiles)
    _clip_pkg = pkg
    return pkg


def build_reference_clip_model(state_dict, device="cpu"):
    """state dict -> reference model through the reference's own build_model (so convert_weights applies), then exactly
    what `load` does (clip_explainability.py:163-168): `.to(device)`, and `.float()` only on CPU — on CUDA the stock
    model stays fp16."""
    import_reference_clip()
    me = importlib.import_module("CLIP.clip.model_explainability")
    sd = {k: v.clone() for k, v in state_dict.items() if not k.startswith("__")}
    model = me.build_model(sd).to(device)
    if str(device) == "cpu":
        model.float()
    return model.eval()


def make_reference_wrapper(model_name: str, state_dict, device="cpu"):
    """A ClipWrapper singleton whose weights come from `state_dict` instead of a download."""
    pkg = import_reference_clip()
    ce = importlib.import_module("CLIP.clip.clip_explainability")
    cg = importlib.import_module("CLIP.clip.clip_gradcam")

    def fake_load(name, device="cpu", **kw):
        model = build_reference_clip_model(state_dict, device)
        return model, ce._transform(model.visual.input_resolution)

    cg.load = fake_load
    pkg.load = fake_load  # the stock second model of ClipWrapper.__init__ (unused by get_clip_saliency)
    for attr in ("clip_model", "clip_preprocess", "clip_gradcam"):
        setattr(pkg.ClipWrapper, attr, None)
    pkg.ClipWrapper(clip_model_type=model_name, device=device)
    return pkg.ClipWrapper


def import_reference_module(name: str):
    """`unet3d`, `net`, `utils`, `train_ovssc`, ... from the reference root."""
    assert reference_available()
    install_shims()
    import_reference_clip()  # net.py imports CLIP.clip
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module(name)
