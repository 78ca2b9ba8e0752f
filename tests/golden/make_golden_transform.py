"""Generate tests/golden/case_ba_transform.npz by running the UNMODIFIED reference: a Morpho_pairwise alignment (vecfld) and
BA_transform (spateo/alignment/transform.py:61-116) on query points, float32 and float64, two deformation scales.
Build-container only:  python tests/golden/make_golden_transform.py"""

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle.ref_harness import load_reference, load_reference_transform  # noqa: E402
from spateo_release_b200.synthetic import make_slice_pair  # noqa: E402

VF_KEYS = ("R", "t", "optimal_R", "optimal_t", "init_R", "init_t", "beta", "Coff", "inducing_variables", "normalize_c")
ND_KEYS = ("mean_transformed", "mean_fixed", "scale_transformed", "scale_fixed")


def main():
    mc, _ = load_reference()
    tr = load_reference_transform()
    out = {}
    for tag, dim in (("2d", 2), ("3d", 3)):
        A, B = make_slice_pair(300, 280, 24, dim=dim, seed=1, warp_amplitude=1.5)
        np.random.seed(0)
        ref = mc.Morpho_pairwise(sampleA=B, sampleB=A, device="cpu", dtype="float32", verbose=False, SVI_mode=False,
                                 max_iter=110, K=15, vecfld_key_added="vf")
        ref.run()
        vf = ref.vecfld
        for k in VF_KEYS:
            out[f"{tag}_vf_{k}"] = np.asarray(vf[k])
        for k in ND_KEYS:
            out[f"{tag}_nd_{k}"] = np.asarray(vf["norm_dict"][k])
        pts = np.random.default_rng(3).uniform(0, 100, size=(57, dim))
        out[f"{tag}_points"] = pts
        for dt in ("float64", "float32"):
            for ds in (1, 0.5):
                X, V, O = tr.BA_transform(vf, pts, deformation_scale=ds, dtype=dt, device="cpu")
                sfx = f"{dt}_{ds}"
                out[f"{tag}_XAHat_{sfx}"], out[f"{tag}_vel_{sfx}"], out[f"{tag}_opt_{sfx}"] = X, V, O
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "case_ba_transform.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
