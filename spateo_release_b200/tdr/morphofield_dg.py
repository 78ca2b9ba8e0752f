"""Differential geometry of the morphometric vector field
(reference: spateo/tdr/morphometrics/morphofield_dg/{GPVectorField,differential_geometry}.py).

The reference evaluates the analytical Jacobian of the Gaussian-process field cell by cell in Python and derives
acceleration / curvature / curl / torsion / divergence in further per-cell loops. Here ONE CUDA kernel
(``spb_field_geometry``, csrc/field.cu) evaluates velocity, Jacobian and every derived quantity per query point in
registers; the Python layer only reproduces the reference's names, argument meaning, output shapes (including its
broadcast quirks for the 3-D curl and the torsion) and where results are stored on the AnnData.

``method == "sparsevfc"`` goes to the third-party ``dynamo`` ``SvcVectorField`` in the reference
(differential_geometry.py:24-28; dynamo-release>=1.4.1 is not vendored: **parity unpinned**). ``SvcVectorField`` below
restates what that class evaluates for a SparseVFC field — v(x) = K(x, X_ctrl) C with the Gaussian kernel and its analytical
Jacobian -2 beta sum_m K_m C_m (x - c_m)^T (dynamo's ``Jacobian_rkhs_gaussian``) — on the same kernel; the derived
quantities are the formulas GPVectorField.py took over from dynamo.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from .. import _capi

_OUTPUTS = ("V", "J", "acc", "acc_mat", "curv", "curv_mat", "curl", "torsion", "div", "det")


def _desc(vf_dict: dict, D: int, nonrigid_only: bool, formula: int) -> "_capi.SpbFieldDesc":
    nd = vf_dict["norm_dict"]
    f = _capi.SpbFieldDesc()
    f.D, f.K = D, int(np.asarray(vf_dict["inducing_variables"]).shape[0])
    f.nonrigid_only, f.curvature_formula = int(bool(nonrigid_only)), int(formula)
    f.beta = float(vf_dict["beta"])
    f.scale_transformed = float(nd["scale_transformed"])
    f.scale_fixed = float(nd["scale_fixed"])
    mt = np.broadcast_to(np.asarray(nd["mean_transformed"], dtype=np.float64).reshape(-1), (D,))
    mf = np.broadcast_to(np.asarray(nd["mean_fixed"], dtype=np.float64).reshape(-1), (D,))
    R = np.asarray(vf_dict.get("R", np.eye(D)), dtype=np.float64).reshape(D, D)
    t = np.asarray(vf_dict.get("t", np.zeros(D)), dtype=np.float64).reshape(-1)
    for d in range(D):
        f.mean_transformed[d], f.mean_fixed[d], f.t[d] = mt[d], mf[d], t[d]
        for e in range(D):
            f.R[d * 3 + e] = R[d, e]
    f.velocity_divisor = 10000.0  # gaussian_process.py:127
    return f


def _is_svc(vf_dict: dict) -> bool:
    return vf_dict.get("method") == "sparsevfc" or ("X_ctrl" in vf_dict and "inducing_variables" not in vf_dict)


def _desc_svc(vf_dict: dict, D: int, formula: int) -> "_capi.SpbFieldDesc":
    """A SparseVFC field v(x) = K(x, X_ctrl) C in raw coordinates: the plain RBF part of the kernel (nonrigid_only with unit
    scales and zero means), velocity not divided."""
    f = _capi.SpbFieldDesc()
    f.D, f.K = D, int(np.asarray(vf_dict["X_ctrl"]).shape[0])
    f.nonrigid_only, f.curvature_formula = 1, int(formula)
    f.beta = float(vf_dict["beta"])
    f.scale_transformed = f.scale_fixed = 1.0
    for d in range(D):
        f.R[d * 3 + d] = 1.0
    f.velocity_divisor = 1.0
    return f


def field_geometry(X: np.ndarray, vf_dict: dict, want=("V", "J"), nonrigid_only: bool = False, formula: int = 2,
                   device=None) -> dict:
    """Evaluate the requested quantities (subset of ``_OUTPUTS``) at raw points ``X`` [n, D]; returns numpy float64."""
    import torch

    _capi.require_cuda()
    lib = _capi.load_library()
    svc = _is_svc(vf_dict)
    if svc and "div_cur_free_kernels" in vf_dict:
        raise NotImplementedError("divergence/curl-free kernels are not implemented")
    if not svc and vf_dict["kernel_type"] != "euc":
        if vf_dict["kernel_type"] == "geodist":
            raise NotImplementedError("geodist is not implemented yet")
        raise ValueError("current only support euc and geodist")
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if X.ndim == 1:
        X = X[None, :]
    n, D = X.shape
    if D not in (2, 3):
        raise ValueError("X has incorrect dimensions.")
    if "torsion" in want and D != 3:
        raise Exception("torsion is only defined in 3 dimension.")
    dev = torch.device("cuda" if device in (None, "cuda") else (f"cuda:{device}" if str(device).isdigit() else device))
    f = _desc_svc(vf_dict, D, formula) if svc else _desc(vf_dict, D, nonrigid_only, formula)
    zk, ck = ("X_ctrl", "C") if svc else ("inducing_variables", "Coff")
    if np.asarray(vf_dict[zk]).shape[1] != D or np.asarray(vf_dict[ck]).shape[1] != D:
        raise ValueError("X has incorrect dimensions.")
    with torch.cuda.device(dev):
        Xd = torch.from_numpy(X).to(dev)
        z = torch.from_numpy(np.ascontiguousarray(vf_dict[zk], dtype=np.float64)).to(dev)
        C = torch.from_numpy(np.ascontiguousarray(vf_dict[ck], dtype=np.float64)).to(dev)
        shapes = {"V": (n, D), "J": (n, D, D), "acc": (n,), "acc_mat": (n, D), "curv": (n,), "curv_mat": (n, D),
                  "curl": (n,) if D == 2 else (n, 3), "torsion": (n, 3), "div": (n,), "det": (n,)}
        bufs = {k: (torch.empty(shapes[k], dtype=torch.float64, device=dev) if k in want else None) for k in _OUTPUTS}
        _capi.check(lib.spb_field_geometry(f, _capi.ptr(Xd), n, _capi.ptr(z), _capi.ptr(C),
                                           *[_capi.ptr(bufs[k]) for k in _OUTPUTS], _capi.current_stream_ptr()),
                    "spb_field_geometry")
        return {k: v.cpu().numpy() for k, v in bufs.items() if v is not None}


def Jacobian_GP_gaussian_kernel(X: np.ndarray, vf_dict: dict, vectorize: bool = False, device=None) -> np.ndarray:
    """GPVectorField.py:143-190 — analytical Jacobian, d-by-d-by-n (d-by-d for a single point)."""
    X = np.asarray(X, dtype=np.float64)
    J = field_geometry(X, vf_dict, want=("J",), device=device)["J"]
    if X.ndim == 1:
        return J[0]
    return np.ascontiguousarray(np.transpose(J, (1, 2, 0)))


class GPVectorField:
    """GPVectorField.py:193-266 — same methods; every ``compute_*`` is one kernel launch over all points."""

    def __init__(self):
        self.data = {}

    def from_adata(self, adata, vf_key: str = "VecFld", nonrigid_only: bool = False):
        if vf_key in adata.uns.keys():
            vf_dict = adata.uns[vf_key]
        else:
            raise Exception(
                f"The {vf_key} that corresponds to the reconstructed vector field is not in ``anndata.uns``."
                f"Please run ``st.align.morpho_align(adata, vecfld_key_added='{vf_key}')`` before running this function."
            )
        self.vf_dict = vf_dict
        self.nonrigid_only = nonrigid_only
        self.func = lambda x: self.compute_velocity(x)
        self.data["X"] = vf_dict["X"]
        self.data["V"] = vf_dict["V"]

    def get_data(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.data["X"], self.data["V"]

    def _need_jacobian(self, **kwargs):
        if self.get_Jacobian(**kwargs) is None:  # what calling the reference's ``None`` Jacobian raises
            raise TypeError("'NoneType' object is not callable")

    def _geom(self, X, want, formula=2):
        return field_geometry(X, self.vf_dict, want=want, nonrigid_only=self.nonrigid_only, formula=formula)

    def compute_velocity(self, X: np.ndarray):
        return self._geom(X, ("V",))["V"]

    def compute_acceleration(self, X: Optional[np.ndarray] = None, **kwargs):
        X = self.data["X"] if X is None else X
        self._need_jacobian(**kwargs)
        g = self._geom(X, ("acc", "acc_mat"))
        return g["acc"], g["acc_mat"]

    def compute_curvature(self, X: Optional[np.ndarray] = None, formula: int = 2, **kwargs):
        X = self.data["X"] if X is None else X
        self._need_jacobian(**kwargs)
        if formula == 2:
            g = self._geom(X, ("curv", "curv_mat"), formula=2)
            return g["curv"], g["curv_mat"]
        if formula == 1:
            return self._geom(X, ("curv",), formula=1)["curv"], None
        # the reference leaves the output at zero for any other formula (GPVectorField.py:44-52)
        return np.zeros(len(X)), None

    def compute_curl(self, X: Optional[np.ndarray] = None, dim1: int = 0, dim2: int = 1, dim3: int = 2,
                     **kwargs) -> np.ndarray:
        X = self.data["X"] if X is None else X
        X = np.asarray(X)
        if dim3 is None or X.shape[1] == 2:
            X = X[:, [dim1, dim2]]
        else:
            X = X[:, [dim1, dim2, dim3]]
        self._need_jacobian(**kwargs)
        curl = self._geom(X, ("curl",))["curl"]
        if X.shape[1] == 3:
            # GPVectorField.py:68-71 assigns the 3-vector into an (n, 3, 3) array: every row repeats the curl vector
            curl = np.ascontiguousarray(np.broadcast_to(curl[:, None, :], (len(X), 3, 3)))
        return curl

    def compute_torsion(self, X: Optional[np.ndarray] = None, **kwargs) -> np.ndarray:
        X = self.data["X"] if X is None else X
        self._need_jacobian(**kwargs)
        tau = self._geom(X, ("torsion",))["torsion"]
        # GPVectorField.py:90-96: the 3-vector is stored into an (n, 3, 3) array (rows repeat)
        return np.ascontiguousarray(np.broadcast_to(tau[:, None, :], (len(tau), 3, 3)))

    def compute_divergence(self, X: Optional[np.ndarray] = None, **kwargs) -> np.ndarray:
        X = self.data["X"] if X is None else X
        kwargs.pop("vectorize_size", None)
        self._need_jacobian(**kwargs)
        return self._geom(X, ("div",))["div"]

    def get_Jacobian(self, method: str = "analytical", **kwargs) -> Callable:
        """GPVectorField.py:251-266 — only the analytical Jacobian exists; rows are d f_i, columns d x_j."""
        if method == "analytical":
            return lambda x: Jacobian_GP_gaussian_kernel(X=x, vf_dict=self.vf_dict)
        return None  # the reference falls through for any other method


class SvcVectorField(GPVectorField):
    """What the reference takes from ``dynamo.vectorfield.scVectorField.SvcVectorField`` (differential_geometry.py:24-28)
    for a field learned by ``morphofield_sparsevfc``: velocity K(x, X_ctrl) C, analytical Jacobian, and the same derived
    quantities (acceleration J v, curvature, curl, torsion, divergence) — one kernel launch each. Parity unpinned (dynamo is
    third-party and absent); checked against the float64 restatement in ``oracle/field_oracle.py``."""

    def from_adata(self, adata, basis=None, vf_key: str = "VecFld", nonrigid_only: bool = False):
        super().from_adata(adata, vf_key=vf_key, nonrigid_only=False)
        if not _is_svc(self.vf_dict):
            raise Exception(f"``anndata.uns[{vf_key}]`` does not hold a sparsevfc field (X_ctrl / C / beta).")

    def get_Jacobian(self, method: str = "analytical", **kwargs) -> Callable:
        if method == "analytical":
            return lambda x: Jacobian_GP_gaussian_kernel(X=x, vf_dict=self.vf_dict)
        raise NotImplementedError("only the analytical Jacobian is available (the numerical one needs numdifftools)")


def _generate_vf_class(adata, vf_key: str, method: str = "gaussian_process", nonrigid_only: bool = False):
    """differential_geometry.py:12-39."""
    if vf_key in adata.uns.keys():
        if method == "gaussian_process":
            vector_field_class = GPVectorField()
            vector_field_class.from_adata(adata, vf_key=vf_key, nonrigid_only=nonrigid_only)
        elif method == "sparsevfc":
            vector_field_class = SvcVectorField()
            vector_field_class.from_adata(adata, basis=None, vf_key=vf_key)
        else:
            raise Exception(
                f"The {method} is not in ``anndata.uns[{vf_key}]``."
                f"Please re-run ``st.tdr.morphofield_gp`` or ``st.tdr.morphofield_sparsevfc`` before running this function."
            )
    else:
        raise Exception(
            f"The {vf_key} that corresponds to the reconstructed vector field is not in ``anndata.uns``."
            f"Please run ``st.align.morpho_align(adata, vecfld_key_added='{vf_key}')`` before running this function."
        )
    return vector_field_class


def _vf(adata, vf_key, nonrigid_only):
    return _generate_vf_class(adata=adata, vf_key=vf_key, method=adata.uns[vf_key]["method"],
                              nonrigid_only=nonrigid_only)


def morphofield_velocity(adata, vf_key: str = "VecFld_morpho", key_added: str = "velocity",
                         nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:42-70 — ``.obsm[key_added]``."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    adata.obsm[key_added] = vfc.func(adata.uns[vf_key]["X"])
    return None if inplace else adata


def morphofield_acceleration(adata, vf_key: str = "VecFld_morpho", key_added: str = "acceleration",
                             method: str = "analytical", nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:73-111 — ``.obs[key_added]`` (norm) and ``.obsm[key_added]`` (vectors)."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    adata.obs[key_added], adata.obsm[key_added] = vfc.compute_acceleration(X=X, method=method)
    return None if inplace else adata


def morphofield_curvature(adata, vf_key: str = "VecFld_morpho", key_added: str = "curvature", formula: int = 2,
                          method: str = "analytical", nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:114-159."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    adata.obs[key_added], adata.obsm[key_added] = vfc.compute_curvature(X=X, formula=formula, method=method)
    return None if inplace else adata


def morphofield_curl(adata, vf_key: str = "VecFld_morpho", key_added: str = "curl", method: str = "analytical",
                     nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:162-204 — magnitude in ``.obs``, vectors in ``.obsm``."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    curl = vfc.compute_curl(X=X, method=method)
    curl_mag = np.sqrt((curl.reshape(len(curl), -1) ** 2).sum(1))
    adata.obs[key_added] = curl_mag
    adata.obsm[key_added] = curl
    return None if inplace else adata


def morphofield_torsion(adata, vf_key: str = "VecFld_morpho", key_added: str = "torsion", method: str = "analytical",
                        nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:207-249 — norm in ``.obs``, matrices in ``.uns``."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    torsion_mat = vfc.compute_torsion(X=X, method=method)
    adata.obs[key_added] = np.sqrt((torsion_mat.reshape(len(torsion_mat), -1) ** 2).sum(1))
    adata.uns[key_added] = torsion_mat
    return None if inplace else adata


def morphofield_divergence(adata, vf_key: str = "VecFld_morpho", key_added: str = "divergence",
                           method: str = "analytical", vectorize_size: Optional[int] = 1000,
                           nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:252-294 (``vectorize_size`` is accepted and irrelevant: one launch covers all cells)."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    adata.obs[key_added] = vfc.compute_divergence(X=X, method=method, vectorize_size=vectorize_size)
    return None if inplace else adata


def morphofield_jacobian(adata, vf_key: str = "VecFld_morpho", key_added: str = "jacobian",
                         method: str = "analytical", nonrigid_only: bool = False, inplace: bool = True):
    """differential_geometry.py:297-341 — determinants in ``.obs``, the d-by-d-by-n tensor in ``.uns``."""
    adata = adata if inplace else adata.copy()
    vfc = _vf(adata, vf_key, nonrigid_only)
    X, V = vfc.get_data()
    vfc.get_Jacobian(method=method)
    g = field_geometry(X, vfc.vf_dict, want=("J", "det"))
    adata.obs[key_added] = list(g["det"])
    adata.uns[key_added] = np.ascontiguousarray(np.transpose(g["J"], (1, 2, 0)))
    return None if inplace else adata
