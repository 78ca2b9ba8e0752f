"""Host-side helpers of the alignment path (validation, dense extraction, normalisation, coarse-init pieces).

These mirror the *behaviour* (names, argument meaning, exceptions) of ``spateo/alignment/methods/utils.py`` and
``spateo/alignment/utils.py`` for the functions the pairwise solver calls; the heavy array math lives in the CUDA library.
All ``file:line`` citations are relative to the reference tree.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Union

import numpy as np
import pandas as pd
from scipy.sparse import issparse


def intersect_lsts(*lsts):
    """Intersection of lists, keeping the order of the first list (reference: unordered set, utils.py:21)."""
    others = [set(l) for l in lsts[1:]]
    return [g for g in lsts[0] if all(g in o for o in others)]


def filter_common_genes(*genes, verbose: bool = True) -> list:
    """utils.py:494-512 — raises ValueError when the samples share no gene."""
    common = intersect_lsts(*[list(g) for g in genes])
    if len(common) == 0:
        raise ValueError("The number of common gene between all samples is 0.")
    return common


def to_dense_matrix(X):
    return X.toarray() if issparse(X) else np.array(X)


def check_spatial_coords(sample, spatial_key: str = "spatial") -> np.ndarray:
    """utils.py:70-108 — returns the coordinates without constant axes; must end up 2-D or 3-D."""
    if spatial_key not in sample.obsm:
        raise KeyError(f"Spatial key '{spatial_key}' not found in AnnData object.")
    coordinates = sample.obsm[spatial_key].copy()
    if isinstance(coordinates, pd.DataFrame):
        coordinates = coordinates.values
    coordinates = np.asarray(coordinates)
    keep = [i for i in range(coordinates.shape[1]) if len(np.unique(coordinates[:, i])) != 1]
    coordinates = coordinates[:, keep]
    if coordinates.shape[1] > 3 or coordinates.shape[1] < 2:
        raise ValueError(f"The spatial coordinate '{spatial_key}' should only has 2 / 3 dimension")
    return np.asarray(coordinates)  # column-major after the fancy column selection, like the reference's (see normalize_coords)


def check_exp(sample, layer: str = "X") -> np.ndarray:
    """utils.py:112-135 — dense expression matrix of ``.X`` or ``.layers[layer]``."""
    if layer == "X":
        m = sample.X
    else:
        if layer not in sample.layers:
            raise KeyError(f"Layer '{layer}' not found in AnnData object.")
        m = sample.layers[layer]
    # The reference copies here (utils.py:127-131) because its backend later normalises in place; this package never
    # writes into a representation (the device holds the working copies), so a dense input is passed through as a view
    # and an 800 MB matrix is not duplicated twice per slice.
    return m.toarray() if issparse(m) else np.asarray(m)


def check_obs(rep_layer: List[str], rep_field: List[str]) -> Optional[str]:
    """utils.py:139-170 — at most one 'obs' (label) representation."""
    pos = [i for i, f in enumerate(rep_field) if f == "obs"]
    if len(pos) > 1:
        raise ValueError("'obs' occurs more than once in the list. Currently Spateo only support one label consistency.")
    return rep_layer[pos[0]] if pos else None


def check_rep_layer(samples, rep_layer="X", rep_field="layer") -> bool:
    """utils.py:174-224 — True, or ValueError naming the missing representation."""
    if isinstance(rep_layer, str):
        rep_layer = [rep_layer]
    if isinstance(rep_field, str):
        rep_field = [rep_field] * len(rep_layer)
    for sample in samples:
        for rep, rep_f in zip(rep_layer, rep_field):
            missing = False
            if rep_f == "layer":
                missing = (rep != "X") and (rep not in sample.layers)
            elif rep_f == "obsm":
                missing = rep not in sample.obsm
            elif rep_f == "obs":
                missing = rep not in sample.obs
                if not missing and not isinstance(sample.obs[rep].dtype, pd.CategoricalDtype):
                    raise ValueError(
                        f"The specified representation '{rep}' found in the '{rep_f}' attribute should be categorical."
                    )
            else:
                raise ValueError("rep_field must be either 'layer', 'obsm' or 'obs'")
            if missing:
                raise ValueError(
                    f"The specified representation '{rep}' not found in the '{rep_f}' attribute of some of the AnnData objects."
                )
    return True


def check_label_transfer_dict(catA, catB, label_transfer_dict):
    """utils.py:228-260 — KeyError when a category pair is missing."""
    for ca in catA:
        if ca not in label_transfer_dict:
            raise KeyError(f"Category '{ca}' from catA not found in label_transfer_dict.")
        for cb in catB:
            if cb not in label_transfer_dict[ca]:
                raise KeyError(
                    f"Category '{cb}' from catB not found in label_transfer_dict for category '{ca}' from catA."
                )


def generate_label_transfer_dict(
    cat1, cat2, positive_pairs=None, negative_pairs=None, default_positive_value: float = 10.0,
    default_negative_value: float = 1.0,
) -> Dict[str, Dict[str, float]]:
    """utils.py:376-437 — row-normalised label transfer prior (same defaults)."""
    table = {c1: {c2: 1.0 for c2 in cat2} for c1 in cat1}
    if positive_pairs is None and negative_pairs is None:
        table = {c1: {c2: default_negative_value for c2 in cat2} for c1 in cat1}
        positive_pairs = [{"left": [c], "right": [c], "value": default_positive_value} for c in np.union1d(cat1, cat2)]
    for pairs in (positive_pairs, negative_pairs):
        if pairs is None:
            continue
        for p in pairs:
            for l in p["left"]:
                for r in p["right"]:
                    if r in table and l in table[r]:
                        table[r][l] = p["value"]
    out = {}
    for c1 in cat1:
        norm = np.array([table[c1][c2] for c2 in cat2]).sum()
        out[c1] = {c2: table[c1][c2] / (norm + 1e-8) for c2 in cat2}
    return out


def check_label_transfer(sampleA, sampleB, obs_key: str, label_transfer_dict=None) -> np.ndarray:
    """utils.py:296-312 — float32 matrix [len(catA), len(catB)]."""
    if label_transfer_dict is not None and not isinstance(label_transfer_dict, dict):
        raise ValueError("label_transfer_dict should be a list or a dictionary.")
    catA = sampleA.obs[obs_key].cat.categories.tolist()
    catB = sampleB.obs[obs_key].cat.categories.tolist()
    if label_transfer_dict is None:
        label_transfer_dict = generate_label_transfer_dict(catA, catB)
    lt = np.zeros((len(catA), len(catB)), dtype=np.float32)
    for j, ca in enumerate(catA):
        for k, cb in enumerate(catB):
            lt[j, k] = label_transfer_dict[ca][cb]
    return lt


def get_rep(sample, rep: str = "X", rep_field: str = "layer", genes=None, dtype=np.float32) -> np.ndarray:
    """utils.py:441-486 — dense host matrix (layer/obsm) or int32 category codes (obs)."""
    if rep_field == "layer":
        return np.ascontiguousarray(check_exp(sample=sample[:, genes], layer=rep), dtype=dtype)
    if rep_field == "obs":
        return np.array(sample.obs[rep].cat.codes.values, dtype=np.int32)
    if rep_field == "obsm":
        return np.ascontiguousarray(np.asarray(sample.obsm[rep]), dtype=dtype)
    raise ValueError("rep_field must be either 'layer', 'obsm' or 'obs'")


def normalize_coords(coordsA: np.ndarray, coordsB: np.ndarray, separate_mean=True, separate_scale=False):
    """morpho_class.py:589-635 — zero-mean per slice, RMS scale (shared by default). Returns new arrays + params.

    The arithmetic follows the reference operation by operation IN THE ARRAYS' OWN DTYPE AND MEMORY LAYOUT (einsum
    reductions of the float32, column-major arrays ``check_spatial_coords`` returns): the coarse initialisation downstream
    sits on a knife edge — ``np.arange(lo, hi, (hi - lo) / n)`` yields n or n + 1 voxel-grid points depending on the last
    bit of ``lo`` / ``hi`` — so a one-ulp difference here changes the voxel set, the inlier pairs and the initial pose."""
    dt = coordsA.dtype
    coords = [coordsA.copy(order="K"), coordsB.copy(order="K")]
    D = coordsA.shape[1]
    means = np.zeros((2, D), dtype=dt)
    scales = np.zeros((2,), dtype=dt)
    for i in range(2):
        means[i] = np.einsum("ij->j", coords[i]) / coords[i].shape[0]
    if not separate_mean:
        means = np.repeat(means.mean(axis=0), 2, axis=0)  # same (odd) behaviour as morpho_class.py:615
    for i in range(2):
        coords[i] -= means[i]
        scales[i] = np.sqrt(np.einsum("ij->", np.einsum("ij,ij->ij", coords[i], coords[i])) / coords[i].shape[0])
    if not separate_scale:
        scales = np.full((2,), scales.mean(), dtype=dt)
    for i in range(2):
        coords[i] /= scales[i]
    return coords[0], coords[1], scales, means


def solve_RT_by_correspondence(X: np.ndarray, Y: np.ndarray, return_scale: bool = False):
    """spateo/alignment/utils.py:350-402 — least-squares R, t with X ~ Y R^T + t (no reflection guard, as there)."""
    tX, tY = np.mean(X, axis=0), np.mean(Y, axis=0)
    Xc, Yc = X - tX, Y - tY
    H = Yc.T @ Xc
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    t = np.mean(Xc, axis=0) - np.mean(Yc, axis=0) + tX - tY @ R.T
    if return_scale:
        s = np.trace(Xc.T @ Xc - R.T @ (Yc.T @ Xc)) / np.trace(Yc.T @ Yc)
        return R, t, s
    return R, t


def empty_cache(device: str = "cpu"):
    """utils.py:1413-1415."""
    import torch

    if torch.cuda.is_available():
        torch.cuda.empty_cache()
