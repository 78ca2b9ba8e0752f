"""Helpers shared by the CPU and GPU field-geometry tests."""

import numpy as np


def load_field(g, tag):
    vf = dict(
        norm_dict={k: g[f"{tag}_nd_{k}"] for k in ("mean_transformed", "scale_transformed", "mean_fixed", "scale_fixed")},
        kernel_type="euc", inducing_variables=g[f"{tag}_inducing_variables"], beta=float(g[f"{tag}_beta"]),
        Coff=g[f"{tag}_Coff"], R=g[f"{tag}_R"], t=g[f"{tag}_t"], method="gaussian_process")
    return vf, g[f"{tag}_X"]


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
