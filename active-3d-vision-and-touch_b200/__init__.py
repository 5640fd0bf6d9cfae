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
"""
import os
import sys

from . import _lib, dist, encoders, graph, host, model, obj_io, ops, policy, recon, utils  # noqa: F401
from .encoders import Mask_Encoder, Positional_Encoder  # noqa: F401
from .model import GCN, Encoder, GCN_layer, Graph_Model  # noqa: F401
from .utils import batch_sample, chamfer_distance  # noqa: F401

__all__ = ["ops", "utils", "model", "graph", "host", "dist", "GCN", "GCN_layer", "batch_sample",
           "chamfer_distance", "install", "install_pytorch3d_shim"]

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pytorch3d_shim")


def install_pytorch3d_shim():
    """Seam S1: make `import pytorch3d.loss` etc. (utils.py:20-23) resolve to the ptk_b200 kernels."""
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)


def install(ref_utils=None, ref_models=()):
    """Seam S2: patch the reference's modules in place so its unchanged scripts use the fused path.

    ref_utils  : the imported `pterotactyl.utility.utils` module (imported here if None)
    ref_models : modules holding GCN / GCN_layer copies (vision.model, autoencoder.model, DDQN.model);
                 vision.model's Positional_Encoder is replaced too
    """
    if ref_utils is None:
        install_pytorch3d_shim()
        from pterotactyl.utility import utils as ref_utils  # noqa: WPS433
    for name in ("chamfer_distance", "batch_sample", "calc_adj", "normalize_adj", "adj_fuse_touch",
                 "adj_init", "load_mesh_touch", "load_mesh_vision"):
        setattr(ref_utils, name, getattr(utils, name))
    for mod in ref_models:
        if hasattr(mod, "GCN_layer"):
            mod.GCN_layer = GCN_layer
        if hasattr(mod, "GCN"):
            mod.GCN = GCN
        if hasattr(mod, "Positional_Encoder"):
            mod.Positional_Encoder = Positional_Encoder
        if hasattr(mod, "Encoder") and hasattr(mod, "AutoEncoder"):  # autoencoder/model.py:45
            mod.Encoder = model.Encoder
        if hasattr(mod, "Graph_Model"):  # policies/DDQN/model.py:65
            mod.Graph_Model = model.Graph_Model
    return ref_utils
