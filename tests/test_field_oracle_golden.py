"""CPU: the numpy restatement of the field geometry (oracle/field_oracle.py) against the fixture generated from the
unmodified reference (tests/golden/make_golden_field.py)."""

import numpy as np
import pytest

from field_helpers import load_field, rel
from oracle import field_oracle as fo


@pytest.mark.parametrize("tag", ["2d", "3d"])
def test_field_oracle_matches_reference(golden, tag):
    g = golden("field_geometry")
    vf, X = load_field(g, tag)
    assert rel(fo.jacobian(X, vf), g[f"{tag}_J"]) < 1e-12
    assert rel(fo.jacobian(X, vf), g[f"{tag}_J_vec"]) < 1e-12
    assert rel(fo.jacobian(X[3:4], vf)[:, :, 0], g[f"{tag}_J_single"]) < 1e-12
    assert rel(fo.curl(X, vf), g[f"{tag}_curl"]) < 1e-12
    assert rel(fo.divergence(X, vf), g[f"{tag}_div"]) < 1e-12
    for nro, sfx in ((False, ""), (True, "_nro")):
        assert rel(fo.gp_velocity(X, vf, nro), g[f"{tag}_V{sfx}"]) < 1e-12
        _, _, acc, acc_mat = fo.acceleration(X, vf, nro)
        assert rel(acc, g[f"{tag}_acc{sfx}"]) < 1e-11 and rel(acc_mat, g[f"{tag}_acc_mat{sfx}"]) < 1e-11
        c2, c2m = fo.curvature(X, vf, 2, nro)
        assert rel(c2, g[f"{tag}_curv2{sfx}"]) < 1e-9 and rel(c2m, g[f"{tag}_curv2_mat{sfx}"]) < 1e-9
        assert rel(fo.curvature(X, vf, 1, nro)[0], g[f"{tag}_curv1{sfx}"]) < 1e-9
        if tag == "3d":
            assert rel(fo.torsion(X, vf, nro), g[f"{tag}_torsion{sfx}"]) < 1e-9
