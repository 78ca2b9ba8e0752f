"""GPU: spb_field_geometry / st.tdr differential-geometry wrappers against the reference fixture (fp64, 1e-9)."""

import numpy as np
import pytest

from field_helpers import load_field, rel

pytestmark = pytest.mark.gpu


class _Adata:
    def __init__(self, n):
        import pandas as pd

        self.uns, self.obsm, self.obs, self.n_obs = {}, {}, pd.DataFrame(index=[str(i) for i in range(n)]), n

    def copy(self):
        import copy

        return copy.deepcopy(self)


@pytest.mark.parametrize("tag", ["2d", "3d"])
def test_field_geometry_matches_reference(golden, tag):
    from spateo_release_b200.tdr import morphofield_dg as dg

    g = golden("field_geometry")
    vf, X = load_field(g, tag)
    assert rel(dg.Jacobian_GP_gaussian_kernel(X, vf), g[f"{tag}_J"]) < 1e-10
    assert rel(dg.Jacobian_GP_gaussian_kernel(X, vf, vectorize=True), g[f"{tag}_J_vec"]) < 1e-10
    Js = dg.Jacobian_GP_gaussian_kernel(X[3], vf)
    assert Js.shape == g[f"{tag}_J_single"].shape and rel(Js, g[f"{tag}_J_single"]) < 1e-10
    for nro, sfx in ((False, ""), (True, "_nro")):
        vf["X"], vf["V"] = X, g[f"{tag}_V{sfx}"]
        a = _Adata(len(X))
        a.uns["VecFld"] = vf
        c = dg.GPVectorField()
        c.from_adata(a, vf_key="VecFld", nonrigid_only=nro)
        assert rel(c.compute_velocity(X), g[f"{tag}_V{sfx}"]) < 1e-10
        acc, acc_mat = c.compute_acceleration()
        assert rel(acc, g[f"{tag}_acc{sfx}"]) < 1e-9 and rel(acc_mat, g[f"{tag}_acc_mat{sfx}"]) < 1e-9
        c2, c2m = c.compute_curvature(formula=2)
        assert rel(c2, g[f"{tag}_curv2{sfx}"]) < 1e-8 and rel(c2m, g[f"{tag}_curv2_mat{sfx}"]) < 1e-8
        c1, none = c.compute_curvature(formula=1)
        assert none is None and rel(c1, g[f"{tag}_curv1{sfx}"]) < 1e-8
        if tag == "3d":
            tor = c.compute_torsion()
            assert tor.shape == g[f"{tag}_torsion{sfx}"].shape and rel(tor, g[f"{tag}_torsion{sfx}"]) < 1e-8
        else:
            with pytest.raises(Exception, match="torsion is only defined"):
                c.compute_torsion()
        if not nro:
            curl = c.compute_curl()
            assert curl.shape == g[f"{tag}_curl"].shape and rel(curl, g[f"{tag}_curl"]) < 1e-10
            assert rel(c.compute_divergence(), g[f"{tag}_div"]) < 1e-10
        with pytest.raises(TypeError):
            c.compute_acceleration(method="numerical")


def test_differential_geometry_wrappers_store_like_reference(golden):
    """differential_geometry.py:42-341 — where each quantity lands on the AnnData, after a real morphofield_gp call."""
    from spateo_release_b200 import tdr
    from spateo_release_b200.tdr import morphofield_dg as dg

    g = golden("field_geometry")
    vf, X = load_field(g, "3d")
    a = _Adata(len(X))
    a.uns["VecFld_morpho"] = vf
    a.obsm["align_spatial"] = X
    tdr.morphofield_gp(a, grid_num=[5, 5, 5])
    assert a.uns["VecFld_morpho"]["method"] == "gaussian_process"
    assert rel(a.uns["VecFld_morpho"]["V"], g["3d_V"]) < 1e-10
    dg.morphofield_velocity(a)
    dg.morphofield_acceleration(a)
    dg.morphofield_curvature(a)
    dg.morphofield_curl(a)
    dg.morphofield_torsion(a)
    dg.morphofield_divergence(a)
    dg.morphofield_jacobian(a)
    assert rel(a.obsm["velocity"], g["3d_V"]) < 1e-10
    assert rel(a.obs["acceleration"].values, g["3d_acc"]) < 1e-9 and rel(a.obsm["acceleration"], g["3d_acc_mat"]) < 1e-9
    assert rel(a.obs["curvature"].values, g["3d_curv2"]) < 1e-8
    assert a.obsm["curl"].shape == (len(X), 3, 3)
    assert rel(a.obs["curl"].values, np.linalg.norm(g["3d_curl"].reshape(len(X), -1), axis=1)) < 1e-10
    assert rel(a.obs["torsion"].values, np.linalg.norm(g["3d_torsion"].reshape(len(X), -1), axis=1)) < 1e-8
    assert rel(a.obs["divergence"].values, g["3d_div"]) < 1e-10
    J = g["3d_J"]
    assert a.uns["jacobian"].shape == J.shape and rel(a.uns["jacobian"], J) < 1e-10
    dets = np.array([np.linalg.det(J[:, :, i]) for i in range(J.shape[2])])
    assert rel(a.obs["jacobian"].values, dets) < 1e-9
    b = dg.morphofield_divergence(a, key_added="div2", inplace=False)
    assert "div2" in b.obs and "div2" not in a.obs


@pytest.mark.parametrize("D", [2, 3])
def test_sparsevfc_field_geometry(D):
    """differential_geometry.py:24-28 — a field learned by morphofield_sparsevfc (dynamo's SvcVectorField in the reference;
    parity unpinned): velocity, Jacobian and every derived quantity against the float64 restatement, through the class and
    through the st.tdr wrappers."""
    from oracle import field_oracle as fo
    from spateo_release_b200.tdr import morphofield_dg as dg

    rng = np.random.default_rng(11 + D)
    n, M = 300, 40
    X = rng.uniform(0, 50, (n, D))
    vf = {"X_ctrl": rng.uniform(0, 50, (M, D)), "C": rng.normal(size=(M, D)), "beta": 1.0 / 12.0**2, "method": "sparsevfc"}
    want = fo.svc_geometry(X, vf)
    vf["X"], vf["V"] = X, want["V"]
    a = _Adata(n)
    a.uns["VecFld_morpho"] = vf
    c = dg._generate_vf_class(a, "VecFld_morpho", method="sparsevfc")
    assert isinstance(c, dg.SvcVectorField)
    assert rel(c.func(X), want["V"]) < 1e-12
    assert rel(c.get_Jacobian()(X), want["J"]) < 1e-11
    assert rel(c.get_Jacobian()(X[2]), want["J"][:, :, 2]) < 1e-11
    acc, acc_mat = c.compute_acceleration()
    assert rel(acc, want["acc"]) < 1e-10 and rel(acc_mat, want["acc_mat"]) < 1e-10
    c2, c2m = c.compute_curvature(formula=2)
    assert rel(c2, want["curv2"]) < 1e-9 and rel(c2m, want["curv2_mat"]) < 1e-9
    assert rel(c.compute_curvature(formula=1)[0], want["curv1"]) < 1e-9
    curl = c.compute_curl()
    assert curl.shape == want["curl"].shape and rel(curl, want["curl"]) < 1e-11
    assert rel(c.compute_divergence(), want["div"]) < 1e-11
    if D == 3:
        assert rel(c.compute_torsion(), want["torsion"]) < 1e-9
    with pytest.raises(NotImplementedError):
        c.compute_acceleration(method="numerical")
    dg.morphofield_velocity(a)
    dg.morphofield_acceleration(a)
    dg.morphofield_divergence(a)
    dg.morphofield_jacobian(a)
    assert rel(a.obsm["velocity"], want["V"]) < 1e-12 and rel(a.obs["acceleration"].values, want["acc"]) < 1e-10
    assert rel(a.obs["divergence"].values, want["div"]) < 1e-11
    assert rel(a.uns["jacobian"], want["J"]) < 1e-11 and rel(a.obs["jacobian"].values, want["det"]) < 1e-9


def test_sparsevfc_field_geometry_after_fit():
    """End to end: st.tdr.morphofield_sparsevfc -> differential geometry of the stored field (keys X_ctrl / C / beta)."""
    from oracle import field_oracle as fo
    from spateo_release_b200 import tdr
    from spateo_release_b200.tdr import morphofield_dg as dg

    rng = np.random.default_rng(3)
    n = 1500
    X = rng.uniform(0, 60, (n, 3))
    c = X - 30.0
    V = np.stack([-0.05 * c[:, 1], 0.05 * c[:, 0], 0.02 * c[:, 2]], axis=1) + rng.normal(0, 0.05, (n, 3))
    a = _Adata(n)
    a.obsm["spatial"], a.obsm["V"] = X, V
    vf = tdr.sparsevfc.SparseVFC(X, V, Grid=None, M=60, lambda_=0.02, MaxIter=30, device="0")
    vf["method"] = "sparsevfc"
    a.uns["VecFld_morpho"] = vf
    want = fo.svc_geometry(vf["X"], {k: np.asarray(vf[k]) for k in ("X_ctrl", "C")} | {"beta": vf["beta"]})
    dg.morphofield_curl(a)
    dg.morphofield_divergence(a)
    assert rel(a.obsm["curl"], want["curl"]) < 1e-9 and rel(a.obs["divergence"].values, want["div"]) < 1e-9
    # the fitted field is a rotation about z (+ stretch along z): curl_z ~ 0.1, divergence ~ 0.02 in the interior
    inner = (np.abs(c) < 15).all(1)
    assert abs(np.median(a.obsm["curl"][inner, 0, 2]) - 0.1) < 0.03
    assert abs(np.median(a.obs["divergence"].values[inner]) - 0.02) < 0.02
