"""Host-buffer entry points (numpy in / numpy out) over ptk_host_* of include/ptk.h: the end-to-end
path a caller without torch binds.  H2D + kernels + D2H happen inside each call."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class HostContext:
    """Owns a CUDA stream and grow-only device buffers on one device."""

    def __init__(self, device=0):
        self._h = _lib.lib().ptk_host_ctx_create(int(device))
        if not self._h:
            raise RuntimeError("ptk_host_ctx_create failed: " + _lib.last_error())

    def close(self):
        if self._h:
            _lib.lib().ptk_host_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chamfer(self, x, y, grad_cham=None, want_idx=False, want_grad_x=True, want_grad_y=True, out=None):
        """x (B,P1,3), y (B,P2,3) float32 host arrays -> dict(cham, [idx_x, idx_y], [grad_x, grad_y])."""
        if x.ndim != 3 or y.ndim != 3 or x.shape[2] != 3 or y.shape[2] != 3 or x.shape[0] != y.shape[0]:
            raise ValueError(f"Expected (B,P,3) point clouds, got {x.shape} and {y.shape}")
        assert x.dtype == np.float32 and y.dtype == np.float32 and x.flags.c_contiguous and y.flags.c_contiguous
        B, P1, _ = x.shape
        P2 = y.shape[1]
        out = out if out is not None else {}
        cham = out.get("cham")
        if cham is None:
            cham = out["cham"] = np.empty(B, np.float32)
        ix = iy = gx = gy = None
        if want_idx:
            ix = out.setdefault("idx_x", np.empty((B, P1), np.int32))
            iy = out.setdefault("idx_y", np.empty((B, P2), np.int32))
        if grad_cham is not None:
            if want_grad_x:
                gx = out.setdefault("grad_x", np.empty((B, P1, 3), np.float32))
            if want_grad_y:
                gy = out.setdefault("grad_y", np.empty((B, P2, 3), np.float32))
        _lib.check(_lib.lib().ptk_host_chamfer(self._h, _p(x), _p(y), B, P1, P2, _p(cham), _p(ix), _p(iy),
                                               _p(grad_cham), _p(gx), _p(gy)), "ptk_host_chamfer")
        return out

    def mesh_chamfer(self, verts, faces, gt, u_face, uv, grad_cd=None):
        """utils.chamfer_distance on host arrays.  u_face (R,B,S), uv (R,2,B,S) -> cd (B,), grad_verts."""
        B, V, _ = verts.shape
        F = faces.shape[0]
        R, _, S = u_face.shape
        P2 = gt.shape[1]
        faces = np.ascontiguousarray(faces, np.int32)
        cd = np.empty(B, np.float32)
        gv = np.empty((B, V, 3), np.float32) if grad_cd is not None else None
        _lib.check(_lib.lib().ptk_host_mesh_chamfer(self._h, _p(verts), B, V, _p(faces), F, _p(gt), P2, _p(u_face),
                                                    _p(uv), S, R, _p(cd), _p(grad_cd), _p(gv)),
                   "ptk_host_mesh_chamfer")
        return cd, gv
