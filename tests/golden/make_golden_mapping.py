"""Generate tests/golden/case_mapping.npz from the UNMODIFIED reference's get_optimal_mapping_relationship /
mapping_aligned_coords (spateo/alignment/utils.py:157-254). Build-container only:
    python tests/golden/make_golden_mapping.py
The pi matrix is a posterior-like matrix with a few all-zero rows and columns so the KD-tree tie-break is exercised."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import load_reference_align_utils  # noqa: E402


def make_inputs(seed=0, nx=90, ny=80, dim=2):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 10, size=(nx, dim))
    Y = rng.uniform(0, 10, size=(ny, dim))
    d2 = ((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
    pi = np.exp(-d2 / 0.5).astype(np.float32)
    pi[pi < 1e-6] = 0.0
    pi[[5, 17, 60], :] = 0.0  # rows / columns whose maximum is shared by every entry
    pi[:, [3, 44]] = 0.0
    pi /= pi.sum(0, keepdims=True) + 1e-8
    return X, Y, pi


def main():
    au = load_reference_align_utils()
    out = {}
    for tag, dim in (("2d", 2), ("3d", 3)):
        X, Y, pi = make_inputs(seed=dim, dim=dim)
        out[f"{tag}_X"], out[f"{tag}_Y"], out[f"{tag}_pi"] = X, Y, pi
        for keep_all in (False, True):
            sfx = "_all" if keep_all else ""
            xi, xv, yi, yv = au.get_optimal_mapping_relationship(X=X, Y=Y, pi=pi, keep_all=keep_all)
            out[f"{tag}_xi{sfx}"], out[f"{tag}_xv{sfx}"], out[f"{tag}_yi{sfx}"], out[f"{tag}_yv{sfx}"] = xi, xv, yi, yv
        mx, my = au.mapping_aligned_coords(X=X, Y=Y, pi=pi, keep_all=False)
        for nm, mp in (("mx", mx), ("my", my)):
            for k, v in mp.items():
                out[f"{tag}_{nm}_{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "case_mapping.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
