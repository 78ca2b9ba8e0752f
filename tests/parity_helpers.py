"""Shared helpers of the GPU parity tests: build the CUDA ``Morpho_pairwise`` from a golden fixture, overwrite its device
state with the reference's E-step inputs, and evaluate the float64 oracle on the same inputs."""

import ast
import ctypes as C

import numpy as np


def adata_from_golden(g):
    import pandas as pd

    from spateo_release_b200.anndata_lite import AnnDataLite

    G = g["exp_moving"].shape[1]
    var = pd.DataFrame(index=[f"g{i}" for i in range(G)])
    mov = AnnDataLite(np.asarray(g["exp_moving"], dtype=np.float32), var=var.copy(), obsm={"spatial": g["raw_coords_moving"]})
    fix = AnnDataLite(np.asarray(g["exp_fixed"], dtype=np.float32), var=var.copy(), obsm={"spatial": g["raw_coords_fixed"]})
    return mov, fix


def cfg_of(g):
    return ast.literal_eval(str(g["cfg"]))


def model_from_golden(g, **over):
    import spateo_release_b200 as st

    cfg = cfg_of(g)
    mov, fix = adata_from_golden(g)
    kw = dict(SVI_mode=cfg["svi"], max_iter=cfg["max_iter"], K=cfg["K"], verbose=False, device="0", vecfld_key_added="vf")
    kw.update(cfg["kw"])
    if "guide_fixed" in g:
        kw["guidance_pair"] = [g["guide_fixed"], g["guide_moving"]]
    kw.update(over)
    np.random.seed(0)
    return st.align.Morpho_pairwise(sampleA=mov, sampleB=fix, **kw)


def relF(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def poke_estep_state(m, XAHat, alpha, SigmaDiag, sigma2, gamma, sigma2_variance, coordsB=None, samples_s=None):
    """Overwrite the device state with given E-step inputs (arrays in the CALLER's row order)."""
    import torch

    s, NA, D = m._state, m.NA, m.D
    dev = m._dev
    XAHat = m._sorted(np.asarray(XAHat, dtype=np.float32))  # device rows are in processing order
    alpha = m._sorted(np.asarray(alpha, dtype=np.float64))
    SigmaDiag = m._sorted(np.asarray(SigmaDiag, dtype=np.float64))
    s["XAHat"][:D, :NA] = torch.from_numpy(np.ascontiguousarray(XAHat.T)).to(dev)
    s["alpha"][:NA] = torch.from_numpy(alpha.astype(np.float32)).to(dev)
    s["SigmaDiag"][:NA] = torch.from_numpy(SigmaDiag.astype(np.float32)).to(dev)
    mmv = alpha * np.exp(-SigmaDiag / sigma2)
    s["mm"][:NA] = torch.from_numpy(mmv.astype(np.float32)).to(dev)
    s["lm"][:NA] = torch.from_numpy(np.log2(mmv).astype(np.float32)).to(dev)
    if coordsB is not None:
        s["xb4"][:, :D] = torch.from_numpy(np.asarray(coordsB, dtype=np.float32)).to(dev)
    sc = m._read_scalars()
    sc.sigma2, sc.gamma, sc.sigma2_variance = float(sigma2), float(gamma), float(sigma2_variance)
    s["sc"].copy_(torch.from_numpy(np.frombuffer(bytes(sc), dtype=np.uint8).copy()))
    if samples_s is not None:
        m._params.samples_s = float(samples_s)


def poke_golden_estep(m, g, it, sfx=""):
    """Device state <- the reference's E-step inputs at iteration ``it`` of a golden fixture."""
    poke_estep_state(
        m, g[f"it{it}_in_XAHat{sfx}"], g[f"it{it}_in_alpha{sfx}"], g[f"it{it}_in_SigmaDiag{sfx}"],
        float(g[f"it{it}_in_sigma2{sfx}"]), float(g[f"it{it}_in_gamma{sfx}"]), float(g[f"it{it}_in_sigma2_variance{sfx}"]),
        coordsB=g["pre_coordsB" + sfx], samples_s=float(g["pre_samples_s" + sfx]),
    )


def run_estep(m, it, cull=None):
    """One E-step on the current device state (through the C ABI); returns the dense posterior in the caller's row order."""
    import torch

    from spateo_release_b200._capi import check, ptr

    if cull is not None:
        m._params.cull = int(bool(cull))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    m._estep_only(it, st)
    Pd = torch.empty((m.NA, m._NBb), dtype=torch.float32, device=m._dev)
    check(m._lib.spb_materialize_P(C.byref(m._params), it, ptr(Pd), m._NBb, st), "materialize")
    torch.cuda.synchronize()
    return m._unsorted(Pd.cpu().numpy())


def device_rows(m, name):
    return m._unsorted(m._state[name][: m.NA].cpu().numpy())


def device_pxb(m):
    return m._unsorted(m._state["PXB"][: m.D, : m.NA].T.contiguous().cpu().numpy())
