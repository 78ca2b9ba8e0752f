"""Duck-typed stand-in for ``anndata.AnnData`` for environments where anndata is not installed.

The alignment path only touches a small surface of AnnData (reference:
spateo/alignment/methods/morpho_class.py:472-484,528-534,928-929 and
spateo/alignment/methods/utils.py:70-135,441-486): ``.X``, ``.layers``, ``.obsm``, ``.obs``, ``.var``
(``.index``, ``.columns``, optional ``highly_variable``), ``.uns``, ``.shape``, ``.obs_names``, ``.copy()`` and
``__getitem__`` for ``adata[0]``, ``adata[idx_array]`` and ``adata[:, gene_names]``.

When the real ``anndata`` package is importable the host code works on real AnnData objects unchanged; this class
exists so tests, ``bench.py`` and the oracle harness can build inputs without it.
"""

from __future__ import annotations

import numpy as np
import pandas as pd


def _slice_rows(x, idx):
    if x is None:
        return None
    return x[idx]


class AnnDataLite:
    def __init__(self, X, obs=None, var=None, obsm=None, layers=None, uns=None):
        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=[f"cell_{i}" for i in range(n)])
        self.var = var if var is not None else pd.DataFrame(index=[f"gene_{i}" for i in range(g)])
        self.obsm = dict(obsm) if obsm is not None else {}
        self.layers = dict(layers) if layers is not None else {}
        self.uns = dict(uns) if uns is not None else {}

    # ---- basic properties -------------------------------------------------
    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def var_names(self):
        return self.var.index

    def uns_keys(self):
        return list(self.uns.keys())

    def obsm_keys(self):
        return list(self.obsm.keys())

    # ---- copying / slicing --------------------------------------------------
    def copy(self, share_X: bool = False):
        """Deep copy like ``AnnData.copy()``. ``share_X=True`` keeps ``.X`` / ``.layers`` by reference (the alignment
        drivers only ever write ``.obsm`` / ``.uns`` of their working copies, so the expression matrices — 800 MB per
        100k-cell slice — need not be duplicated)."""
        import copy as _copy

        if share_X:
            return AnnDataLite(
                X=self.X, obs=self.obs.copy(), var=self.var.copy(),
                obsm={k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsm.items()},
                layers=dict(self.layers), uns=_copy.deepcopy(self.uns),
            )
        return AnnDataLite(
            X=self.X.copy(),
            obs=self.obs.copy(),
            var=self.var.copy(),
            obsm={k: (v.copy() if hasattr(v, "copy") else v) for k, v in self.obsm.items()},
            layers={k: v.copy() for k, v in self.layers.items()},
            uns=_copy.deepcopy(self.uns),
        )

    def _row_index(self, key):
        n = self.shape[0]
        if isinstance(key, (int, np.integer)):
            return np.array([int(key)])
        if isinstance(key, slice):
            return np.arange(n)[key]
        key = np.asarray(key)
        if key.dtype == bool:
            return np.where(key)[0]
        if key.dtype.kind in "iu":
            return key
        # names
        return self.obs.index.get_indexer(key)

    def _col_index(self, key):
        g = self.shape[1]
        if isinstance(key, (int, np.integer)):
            return np.array([int(key)])
        if isinstance(key, slice):
            return np.arange(g)[key]
        if isinstance(key, str):
            key = [key]
        key = np.asarray(list(key)) if not isinstance(key, np.ndarray) else key
        if key.dtype == bool:
            return np.where(key)[0]
        if key.dtype.kind in "iu":
            return key
        idx = self.var.index.get_indexer(key)
        if (idx < 0).any():
            raise KeyError("some variable names are not in var.index")
        return idx

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rkey, ckey = key
        else:
            rkey, ckey = key, slice(None)
        ridx = self._row_index(rkey)
        cidx = self._col_index(ckey)
        n, g = self.shape
        all_rows = len(ridx) == n and np.array_equal(ridx, np.arange(n))
        all_cols = len(cidx) == g and np.array_equal(cidx, np.arange(g))
        if all_rows and all_cols:
            X = self.X  # a view, like anndata's lazy views; consumers copy when they need to
        elif all_rows:
            X = self.X[:, cidx]
        elif all_cols:
            X = self.X[ridx]
        else:
            X = self.X[ridx][:, cidx]
        return AnnDataLite(
            X=X,
            obs=self.obs.iloc[ridx],
            var=self.var.iloc[cidx],
            obsm={k: _slice_rows(np.asarray(v) if not isinstance(v, pd.DataFrame) else v.values, ridx) for k, v in self.obsm.items()},
            layers={k: v[ridx][:, cidx] for k, v in self.layers.items()},
            uns=self.uns,
        )

    def __repr__(self):
        return f"AnnDataLite(n_obs={self.shape[0]}, n_vars={self.shape[1]}, obsm={list(self.obsm)}, layers={list(self.layers)})"


def is_anndata_like(obj) -> bool:
    """True for real AnnData objects and for :class:`AnnDataLite`."""
    return all(hasattr(obj, a) for a in ("X", "obsm", "var", "obs", "uns", "copy"))
