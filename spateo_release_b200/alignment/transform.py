"""``BA_transform`` — apply a learned morpho-align vector field to arbitrary points (spateo/alignment/transform.py:61-116).

The Gaussian-kernel field evaluation ``exp(-beta |q - z|^2) @ Coff`` runs on the device (``spb_field_eval``, fp64); the
3x3 rigid maps are applied on the host.
"""

from __future__ import annotations

import numpy as np
import torch

from .. import _capi
from .._capi import check, ptr


def field_eval(points: np.ndarray, ctrl_pts: np.ndarray, Coff: np.ndarray, beta: float, device=None) -> np.ndarray:
    """``con_K(points, ctrl_pts, beta) @ Coff`` in fp64 on the GPU (utils.py:1132-1158 + transform.py:103)."""
    from .morpho_class import resolve_device

    lib = _capi.load_library()
    dev = resolve_device(device)
    q = np.ascontiguousarray(points, dtype=np.float64)
    z = np.ascontiguousarray(ctrl_pts, dtype=np.float64)
    c = np.ascontiguousarray(Coff, dtype=np.float64)
    n, D = q.shape
    K = z.shape[0]
    if c.ndim == 1:  # vector field never trained (reference keeps Coff = zeros(K), morpho_class.py:733)
        c = np.zeros((K, D))
    with torch.cuda.device(dev):
        qd, zd, cd = (torch.from_numpy(a).to(dev) for a in (q, z, c))
        out = torch.empty((n, D), dtype=torch.float64, device=dev)
        check(lib.spb_field_eval(ptr(qd), n, D, ptr(zd), ptr(cd), K, float(beta), ptr(out), _capi.current_stream_ptr()),
              "spb_field_eval")
        return out.cpu().numpy()


def BA_transform(vecfld, quary_points, deformation_scale: int = 1, dtype: str = "float64", device: str = "cpu"):
    """Apply non-rigid transform to the quary points (same signature and return arity as transform.py:61-67,116).

    Returns ``(XAHat, quary_velocities, quary_optimal_similarity)`` as numpy arrays of ``dtype``.
    """
    dt = np.float32 if dtype == "float32" else np.float64
    f = lambda v: np.asarray(v, dtype=np.float64)
    nd = vecfld["norm_dict"]
    XA = f(quary_points)
    if vecfld["normalize_c"]:
        scale, mean_ref, mean_q = f(nd["scale_transformed"]), f(nd["mean_fixed"]), f(nd["mean_transformed"])
        XA = (XA - mean_q) / scale
    vel = field_eval(XA, vecfld["inducing_variables"], vecfld["Coff"], vecfld["beta"], device) * deformation_scale
    XA = XA @ f(vecfld["init_R"]).T + f(vecfld["init_t"])
    sim = XA @ f(vecfld["R"]).T + f(vecfld["t"])
    opt = XA @ f(vecfld["optimal_R"]).T + f(vecfld["optimal_t"])
    XAHat = vel + sim
    if vecfld["normalize_c"]:
        XAHat = XAHat * scale + mean_ref
        vel = vel * scale
        opt = opt * scale + mean_ref
    return XAHat.astype(dt), vel.astype(dt), opt.astype(dt)
