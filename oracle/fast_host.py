"""TEST INFRASTRUCTURE ONLY — vectorised host restatements of two coarse-initialisation helpers of the reference
(``voxel_data`` spateo/alignment/methods/utils.py:1283-1336, ``inlier_from_NN`` :1220-1280).

They exist so the DEVICE kernels (csrc/voxel.cu, csrc/coarse.cu) can be checked at sizes where the loop restatements in
``oracle/morpho_oracle.py`` (pinned bitwise to the reference) take too long; ``tests/test_host_logic.py`` pins these
against the loop restatements. Like everything under ``oracle/`` they are never imported by the product package.
"""

from typing import Optional

import numpy as np
from scipy.spatial import cKDTree


def voxel_data(coords: np.ndarray, gene_exp: np.ndarray, voxel_size: Optional[float] = None, voxel_num: int = 10000):
    """utils.py:1283-1336, result-identical but without the Python loop over voxels.

    Grid of ``int(sqrt(voxel_num))`` steps per axis; a cell belongs to EVERY grid point closer than ``voxel_size / 2``
    (overlapping membership, not a partition). Candidate pairs come from a KD-tree with a slightly larger radius and are
    then filtered with the reference's own expression evaluated in the input precision, so memberships are identical.
    """
    N, D = coords.shape
    lo, hi = np.min(coords, axis=0), np.max(coords, axis=0)
    if voxel_size is None:
        voxel_size = np.sqrt(np.prod(hi - lo)) / (np.sqrt(N) / 5)
    steps = (hi - lo) / int(np.sqrt(voxel_num))
    axes = [np.arange(a, b, s) for a, b, s in zip(lo, hi, steps)]
    grid = np.stack(np.meshgrid(*axes), axis=-1).reshape(-1, D)
    radius = voxel_size / 2
    tree = cKDTree(np.asarray(coords, dtype=np.float64))
    cand = tree.query_ball_point(np.asarray(grid, dtype=np.float64), r=float(radius) * (1 + 1e-5) + 1e-12)
    counts = np.fromiter((len(c) for c in cand), dtype=np.int64, count=len(cand))
    vox_idx = np.repeat(np.arange(grid.shape[0]), counts)
    cell_idx = np.fromiter((i for c in cand for i in c), dtype=np.int64, count=int(counts.sum()))
    # exact membership test, same arithmetic as the reference (input dtype)
    dist = np.sqrt(np.sum((coords[cell_idx] - grid[vox_idx]) ** 2, axis=1))
    keep = dist < radius
    vox_idx, cell_idx = vox_idx[keep], cell_idx[keep]
    n_in = np.bincount(vox_idx, minlength=grid.shape[0])
    used = n_in > 0
    # voxel means as one sparse membership product (float64 accumulation), only for the non-empty voxels
    from scipy.sparse import csr_matrix

    new_id = np.cumsum(used) - 1
    M = csr_matrix(
        (1.0 / n_in[vox_idx].astype(np.float64), (new_id[vox_idx], cell_idx)), shape=(int(used.sum()), N)
    )
    means = M @ np.asarray(gene_exp, dtype=np.float64)
    return grid[used, :], np.asarray(means)


def inlier_from_NN(train_x, train_y, distance):
    """utils.py:1220-1280 — annealed robust Procrustes on the mutual-NN voxel pairs (host, float64, tiny)."""
    N, D = train_x.shape
    distance = np.maximum(0, distance)
    distance = distance / (np.max(distance) / (np.log(10) * 2))
    alpha, alpha_end, max_iter = 1.0, 0.1, 100
    alpha_dec = np.power(alpha_end / alpha, 1 / (max_iter - 20))
    weight = np.exp(-distance * alpha)
    init_weight = weight
    P = np.ones((N, 1)) * weight
    y_hat = train_x
    sigma2 = np.sum((y_hat - train_y) ** 2) / (D * N)
    gamma = 0.5
    area = np.maximum(np.prod(train_x.max(0) - train_x.min(0)), np.prod(train_y.max(0) - train_y.min(0)))
    Sp = P.sum()
    R, t = np.eye(D), np.zeros(D)
    for it in range(max_iter):
        mu_x = (train_x * P).sum(0) / Sp
        mu_y = (train_y * P).sum(0) / Sp
        A = (train_y - mu_y).T @ ((train_x - mu_x) * P)
        U, _, Vh = np.linalg.svd(A)
        C = np.eye(D)
        C[-1, -1] = np.linalg.det(U @ Vh)
        R = U @ C @ Vh
        t = mu_y - mu_x @ R.T
        y_hat = train_x @ R.T + t
        resid = np.sum((train_y - y_hat) ** 2, 1, keepdims=True)
        term1 = np.exp(-resid / (2 * sigma2)) * weight
        outlier = np.max(weight) * (1 - gamma) * np.power(2 * np.pi * sigma2, D / 2) / (gamma * area)
        P = term1 / (term1 + outlier)
        Sp = P.sum()
        gamma = np.minimum(np.maximum(Sp / N, 0.01), 0.99)
        P = np.maximum(P, 1e-6)
        sigma2 = np.sum((y_hat - train_y) ** 2 * P) / (D * Sp)
        if it > 20:
            alpha = alpha * alpha_dec
            weight = np.exp(-distance * alpha)
            weight = weight / np.max(weight)
    resid = np.sum((train_y - y_hat) ** 2, 1, keepdims=True)
    term1 = np.exp(-resid / (2 * 1e-2)) * weight
    outlier = np.max(weight) * (1 - 0.1) * np.power(2 * np.pi * 1e-2, D / 2) / (0.1 * area)
    P = term1 / (term1 + outlier)
    gamma = np.minimum(np.maximum(P.sum() / N, 0.01), 0.99)
    return P, R, t, init_weight, sigma2, gamma
