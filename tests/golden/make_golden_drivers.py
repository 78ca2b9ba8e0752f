"""Generate tests/golden/case_drivers.npz by running the UNMODIFIED reference drivers
(spateo/alignment/morpho_alignment.py): ``morpho_align`` (serial chain, modes SN-S and SN-N), and
``morpho_align_transformation`` + ``morpho_align_apply_transformation`` (independent pairs on raw coordinates, composed) on
a seeded 4-slice 2-D chain. Build-container only:  python tests/golden/make_golden_drivers.py"""

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")
from oracle.ref_harness import load_reference_drivers  # noqa: E402

from driver_helpers import KW, driver_chain  # noqa: E402


def main():
    drv = load_reference_drivers()
    out = {}
    models, _ = driver_chain()
    for k, m in enumerate(models):
        out[f"in{k}_X"] = np.asarray(m.X, dtype=np.float32)
        out[f"in{k}_spatial"] = np.asarray(m.obsm["spatial"])
    # 1. morpho_align proper: pair i+1 starts from pair i's aligned coordinates (morpho_alignment.py:66-111)
    for mode in ("SN-S", "SN-N"):
        np.random.seed(0)
        aligned, pis = drv.morpho_align([m.copy() for m in models], mode=mode, device="cpu", dtype="float32", verbose=False, **KW)
        tag = mode.replace("-", "")
        for k, a in enumerate(aligned):
            for key in ("align_spatial", "align_spatial_rigid", "align_spatial_nonrigid"):
                out[f"{tag}_{k}_{key}"] = np.asarray(a.obsm[key])
        out[f"{tag}_pi_shapes"] = np.array([p.shape for p in pis])
        out[f"{tag}_pi_sums"] = np.array([float(np.asarray(p, dtype=np.float64).sum()) for p in pis])
        out[f"{tag}_uns_keys_1"] = np.array(sorted(aligned[1].uns.keys()))
    # 2. independent pairs on raw coordinates + composition (morpho_alignment.py:181-217, 284-303)
    np.random.seed(0)
    ms = [m.copy() for m in models]
    tr = drv.morpho_align_transformation(ms, device="cpu", dtype="float32", verbose=False, **KW)
    for i, t in enumerate(tr):
        out[f"tr{i}_Rotation"], out[f"tr{i}_Translation"] = np.asarray(t["Rotation"]), np.asarray(t["Translation"])
    placed = drv.morpho_align_apply_transformation(ms, transformation=tr, verbose=False)
    for k, a in enumerate(placed):
        out[f"placed{k}"] = np.asarray(a.obsm["align_spatial"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "case_drivers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
