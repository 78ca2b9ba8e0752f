"""Morphometric vector fields (reference: spateo/tdr/morphometrics/morphofield/{gaussian_process,sparsevfc}.py).

``morphofield_gp`` evaluates the Gaussian-process field learned by ``st.align.morpho_align`` on cells and grid points;
``morphofield_sparsevfc`` (alias ``morphofield``, the name the tutorials use) learns a sparse kernel vector field from
per-cell displacement vectors. Both share the device RBF kernels of the alignment path.
"""

from __future__ import annotations

from typing import List, Optional, Tuple, Union

import numpy as np

from ..alignment.transform import field_eval


def _grid_from_points(X: np.ndarray, grid_num: List[int]) -> np.ndarray:
    """Bounding-box grid with 1 % margins, as ``get_X_Y_grid`` builds it (spateo/tdr/interpolations/utils.py:39-47)."""
    lo, hi = X.min(0), X.max(0)
    lo = lo - 0.01 * np.abs(hi - lo)
    hi = hi + 0.01 * np.abs(hi - lo)
    axes = np.meshgrid(*[np.linspace(a, b, k) for a, b, k in zip(lo, hi, grid_num)])
    return np.array([a.flatten() for a in axes]).T


def _gp_velocity(X: np.ndarray, vf_dict: dict, nonrigid_only: bool = False, device=None) -> np.ndarray:
    """gaussian_process.py:102-127 — (x_new - x) / 10000 with the field evaluated on the GPU."""
    nd = vf_dict["norm_dict"]
    norm_x = (X - nd["mean_transformed"]) / nd["scale_transformed"]
    if vf_dict["kernel_type"] == "euc":
        vel = field_eval(norm_x, vf_dict["inducing_variables"], vf_dict["Coff"], vf_dict["beta"], device)
    elif vf_dict["kernel_type"] == "geodist":
        raise NotImplementedError("geodist is not implemented yet")
    else:
        raise ValueError("current only support cdist and geodist")
    if nonrigid_only:
        out = vel * nd["scale_fixed"] + (nd["scale_fixed"] - nd["scale_transformed"]) * norm_x
    else:
        rigid = norm_x @ np.asarray(vf_dict["R"], dtype=np.float64).T + np.asarray(vf_dict["t"], dtype=np.float64)
        out = (vel + rigid) * nd["scale_fixed"] + nd["mean_fixed"] - X
    return out / 10000


def morphofield_gp(
    adata,
    spatial_key: str = "align_spatial",
    vf_key: str = "VecFld_morpho",
    NX: Optional[np.ndarray] = None,
    grid_num: Optional[List[int]] = None,
    nonrigid_only: bool = False,
    inplace: bool = True,
    device=None,
):
    """gaussian_process.py:173-233 — fills ``adata.uns[vf_key]`` with X, V, grid, grid_V, method."""
    adata = adata if inplace else adata.copy()
    if vf_key not in adata.uns.keys():
        raise Exception(
            f"The {vf_key} that corresponds to the reconstructed vector field is not in ``anndata.uns``."
            f"Please run ``st.align.morpho_align(adata, vecfld_key_added='{vf_key}')`` before running this function."
        )
    vf = adata.uns[vf_key]
    vf["X"] = np.asarray(adata.obsm[spatial_key], dtype=float)
    vf["V"] = _gp_velocity(vf["X"], vf, nonrigid_only, device)
    if NX is not None:
        predict_X = NX
    else:
        if grid_num is None:
            grid_num = [50, 50, 50]
        predict_X = _grid_from_points(vf["X"], grid_num[: vf["X"].shape[1]])
    vf["grid"] = predict_X
    vf["grid_V"] = _gp_velocity(np.asarray(predict_X, dtype=float), vf, nonrigid_only, device)
    vf["method"] = "gaussian_process"
    return None if inplace else adata


def morphofield_sparsevfc(
    adata,
    spatial_key: str = "align_spatial",
    V_key: str = "V_mapping",
    key_added: str = "VecFld_morpho",
    NX: Optional[np.ndarray] = None,
    grid_num: Optional[List[int]] = None,
    M: int = 100,
    lambda_: float = 0.02,
    lstsq_method: str = "scipy",
    min_vel_corr: float = 0.8,
    restart_num: int = 10,
    restart_seed: Union[List[int], Tuple[int], np.ndarray] = (0, 100, 200, 300, 400),
    inplace: bool = True,
    **kwargs,
):
    """sparsevfc.py:241-328 — SparseVFC vector field with restarts. The solver replaces third-party
    ``dynamo.vectorfield.scVectorField.SparseVFC`` (not vendored by the reference): parity unpinned."""
    from .sparsevfc import morphofield_sparsevfc_core

    adata = adata if inplace else adata.copy()
    adata.uns[key_added] = morphofield_sparsevfc_core(
        X=np.asarray(adata.obsm[spatial_key], dtype=float), V=np.asarray(adata.obsm[V_key], dtype=float), NX=NX,
        grid_num=grid_num, M=M, lambda_=lambda_, lstsq_method=lstsq_method, min_vel_corr=min_vel_corr,
        restart_num=restart_num, restart_seed=restart_seed, **kwargs,
    )
    return None if inplace else adata


# name used by the tutorials (docs/tutorials/notebooks/7_morphogenesis/2_morphogenesis.ipynb)
morphofield = morphofield_sparsevfc
