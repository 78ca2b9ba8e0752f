"""TEST INFRASTRUCTURE ONLY — numpy restatement of spateo/alignment/utils.py:157-191 (get_optimal_mapping_relationship)
and :194-254 (mapping_aligned_coords). Pinned against the unmodified reference by tests/golden/make_golden_mapping.py
(fixture tests/golden/case_mapping.npz). Only tests may import this module."""

import numpy as np
from scipy.spatial import cKDTree


def _resolve(index_pairs, key_col, other_col, P_from, P_to):
    """Keep one (i, j) per distinct value of column ``key_col``; repeated keys are resolved by the nearest coordinate of
    the candidates to the key's own point (utils.py:166-185)."""
    values, counts = np.unique(index_pairs[:, key_col], return_counts=True)
    uniq, rep = values[counts == 1], values[counts != 1]
    out = index_pairs[np.isin(index_pairs[:, key_col], uniq)]
    for i in rep:
        cand = index_pairs[index_pairs[:, key_col] == i]
        _, ii = cKDTree(P_to[cand[:, other_col]]).query(P_from[i], k=1)
        out = np.concatenate([out, cand[ii].reshape(1, 2)], axis=0)
    return out


def get_optimal_mapping_relationship(X, Y, pi, keep_all=False):
    X_max_index = np.argwhere((pi.T == pi.T.max(axis=0)).T)
    Y_max_index = np.argwhere(pi == pi.max(axis=0))
    if not keep_all:
        X_max_index = _resolve(X_max_index, 0, 1, X, Y)
        Y_max_index = _resolve(Y_max_index, 1, 0, Y, X)
    X_pi_value = pi[X_max_index[:, 0], X_max_index[:, 1]].reshape(-1, 1)
    Y_pi_value = pi[Y_max_index[:, 0], Y_max_index[:, 1]].reshape(-1, 1)
    return X_max_index, X_pi_value, Y_max_index, Y_pi_value
