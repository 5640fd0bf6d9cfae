"""ptk_b200 -- the B200 (sm_100a) reconstruction hot path of pterotactyl
(facebookresearch/Active-3D-Vision-and-Touch): batched Chamfer / 1-NN, area-weighted surface
sampling and the mesh-deformation GCN (aggregation + per-vertex linear layers, fwd + bwd).

The directory is named `active-3d-vision-and-touch_b200`; it is imported as `ptk_b200` through the
loader module /ptk_b200.py at the repo root.

Layout
    csrc/ + include/ptk.h   hand-written CUDA kernels behind a C ABI (libptk_b200.so)
    _lib.py                 ctypes binding (fails loudly when the extension is missing)
    ops.py                  torch.autograd bindings (device pointers + current stream)
    utils.py, model.py      the reference's own signatures (utility.utils, GCN_layer, GCN)
    encoders.py             Positional_Encoder / Mask_Encoder with the fused NeRF embedding
    host.py                 host-buffer (numpy) entry points = the non-torch end-to-end path
    recon.py                the GCN + Chamfer-loss part of one reconstruction step (Deformation's layout)
    policy.py               batched candidate scoring of the greedy touch policies (environment.best_step)
    dist.py                 object/batch sharding + NCCL gradient all-reduce / loss gather
    pytorch3d_shim/         the five pytorch3d.* names the reference imports
    import_stubs.py         import-only placeholders for matplotlib / trimesh / pyrender / pybullet / submitit
"""
import os
import sys

from . import _lib, dist, encoders, graph, host, import_stubs, model, obj_io, ops, policy, recon, utils  # noqa: F401
from .import_stubs import install_import_stubs  # noqa: F401
from .encoders import Mask_Encoder, Positional_Encoder  # noqa: F401
from .model import GCN, Encoder, GCN_layer, Graph_Model  # noqa: F401
from .utils import batch_sample, chamfer_distance  # noqa: F401

__all__ = ["ops", "utils", "model", "graph", "host", "dist", "GCN", "GCN_layer", "batch_sample",
           "chamfer_distance", "install", "install_pytorch3d_shim", "install_import_stubs", "find_reference"]

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pytorch3d_shim")


def install_pytorch3d_shim():
    """Seam S1: make `import pytorch3d.loss` etc. (utils.py:20-23) resolve to the ptk_b200 kernels."""
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)


def find_reference():
    """Directory that holds an importable `pterotactyl/` (the reference checkout or its install), or None.
    Order: $PTK_REFERENCE, an already importable `pterotactyl`, <repo>/baseline/_ref (tools/install_reference.py),
    /root/reference."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("PTK_REFERENCE"), None, os.path.join(root, "baseline", "_ref"), "/root/reference"]
    for c in cands:
        if c is None:
            try:
                spec = importlib.util.find_spec("pterotactyl")
            except (ImportError, ValueError):
                spec = None
            if spec is not None and spec.submodule_search_locations:
                return os.path.dirname(list(spec.submodule_search_locations)[0])
            continue
        if os.path.isfile(os.path.join(c, "pterotactyl", "utility", "utils.py")):
            return c
    return None


def import_reference(*names):
    """Import unedited reference modules (`"utility.utils"`, `"reconstruction.vision.model"`, ...) with the
    pytorch3d shim and the import stubs in place; returns them in order.  Raises ImportError when no reference
    checkout / install can be found (find_reference)."""
    import importlib
    install_pytorch3d_shim()
    install_import_stubs()
    ref = find_reference()
    if ref is None:
        raise ImportError("no pterotactyl checkout found: set PTK_REFERENCE or run tools/install_reference.py")
    if ref not in sys.path:
        sys.path.append(ref)
    mods = [importlib.import_module("pterotactyl." + n) for n in names]
    return mods[0] if len(mods) == 1 else mods


_originals = []  # (module, attribute, value before install()) -- what uninstall() restores


def _patch(mod, name, value):
    _originals.append((mod, name, getattr(mod, name, None)))
    setattr(mod, name, value)


def install(ref_utils=None, ref_models=(), face_draw="multinomial"):
    """Seam S2: patch the reference's modules in place so its unchanged scripts use the fused path.

    ref_utils  : the imported `pterotactyl.utility.utils` module (imported here, through the pytorch3d shim and
                 the import stubs, if None)
    ref_models : modules holding GCN / GCN_layer copies (vision.model, autoencoder.model, DDQN.model);
                 their Positional_Encoder copies are replaced too
    face_draw  : how the patched batch_sample / chamfer_distance draw faces (ptk_b200.utils.face_draw).  The default
                 "multinomial" keeps the reference's RNG stream (ATen's Tensor.multinomial does the draw, utils.py:170):
                 with the same seed the scripts sample exactly the points they sampled before.  "uniform" is the fully
                 fused explicit-uniform draw.
    `uninstall()` puts every replaced attribute (and the draw mode) back.
    """
    if ref_utils is None:
        ref_utils = import_reference("utility.utils")
    _patch(utils, "face_draw", face_draw)
    for name in ("chamfer_distance", "batch_sample", "calc_adj", "normalize_adj", "adj_fuse_touch",
                 "adj_init", "load_mesh_touch", "load_mesh_vision"):
        _patch(ref_utils, name, getattr(utils, name))
    for mod in ref_models:
        if hasattr(mod, "GCN_layer"):
            _patch(mod, "GCN_layer", GCN_layer)
        if hasattr(mod, "GCN"):
            _patch(mod, "GCN", GCN)
        if hasattr(mod, "Positional_Encoder"):
            _patch(mod, "Positional_Encoder", Positional_Encoder)
        if hasattr(mod, "Encoder") and hasattr(mod, "AutoEncoder"):  # autoencoder/model.py:45
            _patch(mod, "Encoder", model.Encoder)
        if hasattr(mod, "Graph_Model"):  # policies/DDQN/model.py:65
            _patch(mod, "Graph_Model", model.Graph_Model)
    return ref_utils


def uninstall():
    """Undo every install() since the last uninstall(): the reference's own functions and classes are back."""
    while _originals:
        mod, name, value = _originals.pop()
        if value is None:
            try:
                delattr(mod, name)
            except AttributeError:
                pass
        else:
            setattr(mod, name, value)
