"""Seeded synthetic slice pairs used by the tests, the golden-fixture generator and ``bench.py``.

coords ~ U(0,100)^D; expression = Poisson(exp(sin(coords @ W / 30 + phi))) with W ~ N(0,1)^{D x G}, phi ~ U(0, 2 pi)
(spatially smooth programmes so the KL cost is informative); slice B = slice A rotated by ``theta`` about z, translated,
jittered, optionally warped by a smooth non-rigid field, **row-permuted**, with independently re-sampled counts.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from .anndata_lite import AnnDataLite


def _rotation(D: int, theta: float) -> np.ndarray:
    R = np.eye(D)
    c, s = np.cos(theta), np.sin(theta)
    R[0, 0], R[0, 1], R[1, 0], R[1, 1] = c, -s, s, c
    return R


def _warp(coords: np.ndarray, rng: np.random.Generator, amplitude: float) -> np.ndarray:
    D = coords.shape[1]
    out = coords.copy()
    centres = rng.uniform(20, 80, size=(4, D))
    dirs = rng.normal(size=(4, D))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    for c, d in zip(centres, dirs):
        w = np.exp(-np.sum((coords - c) ** 2, axis=1) / (2 * 20.0**2))
        out += amplitude * w[:, None] * d[None, :]
    return out


def make_slice_pair(
    n_a: int = 1000,
    n_b: int | None = None,
    n_genes: int = 100,
    dim: int = 2,
    seed: int = 0,
    theta: float = 0.5,
    shift: float = 5.0,
    jitter: float = 0.3,
    warp_amplitude: float = 0.0,
    z_thickness: float | None = None,
    dtype=np.float32,
    as_anndata: bool = True,
):
    """Return ``(slice_A, slice_B)``: B is a transformed, permuted, re-sampled copy of (a subset of) A."""
    rng = np.random.default_rng(seed)
    n_b = n_a if n_b is None else n_b
    n = max(n_a, n_b)
    coords = rng.uniform(0, 100, size=(n, dim))
    if dim == 3 and z_thickness is not None:
        coords[:, 2] = rng.uniform(0, z_thickness, size=n)
    W = rng.normal(size=(dim, n_genes))
    phi = rng.uniform(0, 2 * np.pi, size=(n_genes,))

    def counts(c):
        lam = np.exp(np.sin(c @ W / 30.0 + phi))
        return rng.poisson(lam).astype(dtype)

    idx_a = np.arange(n)[:n_a]
    idx_b = rng.permutation(n)[:n_b]
    coords_a = coords[idx_a]
    exp_a = counts(coords_a)
    base_b = coords[idx_b]
    exp_b = counts(base_b)
    moved = base_b
    if warp_amplitude > 0:
        moved = _warp(moved, rng, warp_amplitude)
    R = _rotation(dim, theta)
    coords_b = moved @ R.T + shift + rng.normal(0, jitter, size=moved.shape)

    if not as_anndata:
        return (coords_a, exp_a), (coords_b, exp_b)
    var = pd.DataFrame(index=[f"g{i}" for i in range(n_genes)])
    A = AnnDataLite(exp_a, obs=pd.DataFrame(index=[f"a{i}" for i in range(n_a)]), var=var.copy(), obsm={"spatial": coords_a})
    B = AnnDataLite(exp_b, obs=pd.DataFrame(index=[f"b{i}" for i in range(n_b)]), var=var.copy(), obsm={"spatial": coords_b})
    return A, B


def make_slice_chain(n_slices: int, n_cells: int, n_genes: int, seed: int = 0, theta_step: float = 0.15, dim: int = 2):
    """A serial chain of slices (config 3/4): slice k+1 is slice k rotated by ``theta_step`` and shifted."""
    slices = []
    A, B = make_slice_pair(n_cells, n_cells, n_genes, dim=dim, seed=seed, theta=theta_step)
    slices.append(A)
    slices.append(B)
    for k in range(2, n_slices):
        _, nxt = make_slice_pair(n_cells, n_cells, n_genes, dim=dim, seed=seed + k, theta=theta_step * k)
        slices.append(nxt)
    return slices
