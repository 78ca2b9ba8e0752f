"""``st.tdr.kernel_interpolation`` (spateo/tdr/interpolations/interpolation_sparseVFC.py:13-100): learn a continuous map from
space to expression / label values with the SparseVFC kernel regression and evaluate it on new points.

Same signature, argument meaning and output object as the reference; the regression runs on the device
(``tdr.sparsevfc.SparseVFC``, general output dimension). SparseVFC itself is third-party in the reference (dynamo):
**parity unpinned**, checked against the float64 numpy restatement ``oracle.morpho_oracle.sparse_vfc``.
"""

from __future__ import annotations

from typing import Optional, Union

import numpy as np
import pandas as pd
from scipy.sparse import issparse

from .sparsevfc import SparseVFC


def kernel_interpolation(
    source_adata,
    target_points: Optional[np.ndarray] = None,
    keys: Union[str, list] = None,
    spatial_key: str = "spatial",
    layer: str = "X",
    lambda_: float = 0.02,
    lstsq_method: str = "scipy",
    **kwargs,
):
    """Returns an AnnData-like object holding the interpolated ``keys`` (``.obs`` columns for keys found in
    ``source_adata.obs``, ``.X`` / ``.var`` for gene names) at ``target_points`` (``.obsm[spatial_key]``)."""
    source = source_adata.copy()
    X_src = source.X if layer == "X" else source.layers[layer]
    coords = np.asarray(source.obsm[spatial_key])
    assert keys is not None, "`keys` cannot be None."
    keys = [keys] if isinstance(keys, str) else list(keys)
    obs_keys = [k for k in keys if k in source.obs.keys()]
    var_names = list(source.var_names)
    var_keys = [k for k in keys if k in var_names]
    blocks = []
    if obs_keys:
        blocks.append(np.asarray(source.obs[obs_keys].values, dtype=np.float64))
    if var_keys:
        cols = [var_names.index(k) for k in var_keys]
        sub = X_src[:, cols]
        blocks.append(np.asarray(sub.toarray() if issparse(sub) else sub, dtype=np.float64))
    if not blocks:
        raise ValueError("none of `keys` was found in `.obs` or `.var_names`")
    info = np.concatenate(blocks, axis=1)
    res = SparseVFC(coords, info, target_points, lambda_=lambda_, lstsq_method=lstsq_method, **kwargs)
    target_info = res["grid_V"]
    obs = pd.DataFrame(target_info[:, : len(obs_keys)], columns=obs_keys) if obs_keys else None
    Xout = target_info[:, len(obs_keys):] if var_keys else np.zeros((target_info.shape[0], 0))
    var = pd.DataFrame(index=var_keys) if var_keys else None
    try:
        from anndata import AnnData as _AnnData
    except ImportError:  # the light stand-in carries the same fields
        from ..anndata_lite import AnnDataLite as _AnnData
    return _AnnData(X=Xout, obs=obs, var=var, obsm={spatial_key: np.asarray(target_points)})
