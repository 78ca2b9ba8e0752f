"""AnnData-level drivers with the reference signatures (spateo/alignment/morpho_alignment.py:22-454)."""

from __future__ import annotations

import os
import shutil
from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np

from ..anndata_lite import is_anndata_like
from .morpho_class import Morpho_pairwise
from .transform import BA_transform
from .utils import empty_cache, solve_RT_by_correspondence


def _read_h5ad(path):
    try:
        import anndata as ad
    except ImportError as e:  # pragma: no cover - anndata is optional in this image
        raise ImportError("reading .h5ad files needs the `anndata` package") from e
    return ad.read_h5ad(path)


def morpho_align(
    models: List,
    rep_layer: Union[str, List[str]] = "X",
    rep_field: Union[str, List[str]] = "layer",
    genes: Optional[Union[List[str], np.ndarray]] = None,
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    iter_key_added: Optional[str] = "iter_spatial",
    vecfld_key_added: str = "VecFld_morpho",
    mode: str = "SN-S",
    dissimilarity: Union[str, List[str]] = "kl",
    max_iter: int = 200,
    dtype: str = "float32",
    device: str = "cpu",
    verbose: bool = True,
    **kwargs,
) -> Tuple[List, List[np.ndarray]]:
    """Serial alignment of consecutive slices; pair i+1 starts from pair i's aligned coordinates
    (morpho_alignment.py:22-111). Returns ``(align_models, pis)`` with ``pis[i] = P.T``."""
    align_models = [model.copy() for model in models]
    for m in align_models:
        m.obsm[key_added] = m.obsm[spatial_key].copy()
        m.obsm[f"{key_added}_rigid"] = m.obsm[spatial_key].copy()
        m.obsm[f"{key_added}_nonrigid"] = m.obsm[spatial_key].copy()
    pis = []
    for i in range(len(align_models) - 1):
        modelA, modelB = align_models[i], align_models[i + 1]
        morpho_model = Morpho_pairwise(
            sampleA=modelB,  # moving
            sampleB=modelA,  # fixed
            rep_layer=rep_layer, rep_field=rep_field, dissimilarity=dissimilarity, genes=genes, spatial_key=key_added,
            key_added=key_added, iter_key_added=iter_key_added, vecfld_key_added=vecfld_key_added, max_iter=max_iter,
            dtype=dtype, device=device, verbose=verbose, **kwargs,
        )
        P = morpho_model.run()
        modelB.obsm[f"{key_added}_rigid"] = morpho_model.optimal_RnA.copy()
        modelB.obsm[f"{key_added}_nonrigid"] = morpho_model.XAHat.copy()
        if mode == "SN-S":
            modelB.obsm[key_added] = modelB.obsm[f"{key_added}_rigid"]
        elif mode == "SN-N":
            modelB.obsm[key_added] = modelB.obsm[f"{key_added}_nonrigid"]
        if iter_key_added is not None:
            modelB.uns[iter_key_added] = morpho_model.iter_added
        if vecfld_key_added is not None:
            modelB.uns[vecfld_key_added] = morpho_model.vecfld
        pis.append(P.T if P is not None else None)
        del morpho_model
        empty_cache(device=device)
    return align_models, pis


def pair_transformation(modelA, modelB, spatial_key="spatial", **pairwise_kwargs) -> dict:
    """One link of the chain: align ``modelB`` (moving) onto ``modelA`` (fixed) on RAW coordinates and return the 2-D
    similarity that maps B's raw coordinates onto the aligned ones (morpho_alignment.py:189-211)."""
    pairwise_kwargs.setdefault("materialize_P", False)
    model = Morpho_pairwise(sampleA=modelB, sampleB=modelA, spatial_key=spatial_key, **pairwise_kwargs)
    model.run()
    R, t = solve_RT_by_correspondence(model.optimal_RnA[:, :2], np.asarray(modelB.obsm[spatial_key])[:, :2])
    return {"Rotation": R, "Translation": t}


def morpho_align_transformation(
    models: List,
    models_path: Optional[str] = None,
    save_transformation: bool = False,
    transformation_path: Optional[str] = "./Spateo_transformation",
    resume: bool = False,
    rep_layer: Union[str, List[str]] = "X",
    rep_field: Union[str, List[str]] = "layer",
    genes: Optional[Union[List[str], np.ndarray]] = None,
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    iter_key_added: Optional[str] = "iter_spatial",
    vecfld_key_added: str = "VecFld_morpho",
    dissimilarity: Union[str, List[str]] = "kl",
    max_iter: int = 200,
    dtype: str = "float32",
    device: str = "cpu",
    verbose: bool = True,
    **kwargs,
):
    """Independent pairwise alignments on raw coordinates -> list of {"Rotation", "Translation"} with optional
    per-pair ``.npy`` checkpoints and resume (morpho_alignment.py:114-218)."""
    if models_path is not None:
        assert all(isinstance(m, str) for m in models), "models should be a list of file name if models_path is given."
        assert all(os.path.exists(os.path.join(models_path, m)) for m in models), "Some files in models_path do not exist."
    else:
        assert all(is_anndata_like(m) for m in models), "models should be a list of anndata if models_path is not given."
    iteration, transformation = 0, []
    if save_transformation:
        Path(transformation_path).mkdir(parents=True, exist_ok=True)
        if resume:
            for i in range(len(models) - 1):
                f = os.path.join(transformation_path, f"transformation_{i}.npy")
                if os.path.exists(f):
                    iteration = i
                    transformation.append(np.load(f, allow_pickle=True))
        else:
            remove_all_files_in_directory(transformation_path)
    if models_path is not None:
        modelA = _read_h5ad(os.path.join(models_path, models[iteration]))
    for i in range(iteration, len(models) - 1):
        if models_path is not None:
            modelB = _read_h5ad(os.path.join(models_path, models[i + 1]))
        else:
            modelA, modelB = models[i], models[i + 1]
        cur = pair_transformation(
            modelA, modelB, spatial_key=spatial_key, rep_layer=rep_layer, rep_field=rep_field,
            dissimilarity=dissimilarity, genes=genes, key_added=key_added, iter_key_added=iter_key_added,
            vecfld_key_added=vecfld_key_added, max_iter=max_iter, dtype=dtype, device=device, verbose=verbose, **kwargs,
        )
        transformation.append(cur)
        if save_transformation:
            np.save(os.path.join(transformation_path, f"transformation_{i}.npy"), cur)
        if models_path is not None:
            modelA = modelB
    return transformation


def compose_transformations(transformation: List[dict]):
    """Serial prefix composition of per-pair similarities (morpho_alignment.py:274-301):
    ``cur_t = t_i @ cur_R.T + cur_t ; cur_R = cur_R @ R_i``. Returns the cumulative (R, t) of slices 1..n-1."""
    cur_R, cur_t = np.diag((1.0, 1.0)), np.zeros((2,))
    out = []
    for tr in transformation:
        tr = tr.item() if isinstance(tr, np.ndarray) and tr.dtype == object else tr
        cur_t = tr["Translation"] @ cur_R.T + cur_t
        cur_R = cur_R @ tr["Rotation"]
        out.append((cur_R.copy(), cur_t.copy()))
    return out


def morpho_align_apply_transformation(
    models: List,
    models_path: Optional[str] = None,
    transformation: List[dict] = None,
    transformation_path: Optional[str] = "./Spateo_transformation",
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    save_models_path: Optional[str] = None,
    verbose: bool = True,
):
    """Apply the composed chain of 2-D similarities to every slice (morpho_alignment.py:221-314)."""
    if models_path is not None:
        assert all(isinstance(m, str) for m in models), "models should be a list of file name if models_path is given."
        assert all(os.path.exists(os.path.join(models_path, m)) for m in models), "Some files in models_path do not exist."
    else:
        assert all(is_anndata_like(m) for m in models), "models should be a list of anndata if models_path is not given."
    if transformation is None:
        assert os.path.exists(transformation_path), "transformation_path does not exist."
        transformation = [
            np.load(os.path.join(transformation_path, f"transformation_{i}.npy"), allow_pickle=True)
            for i in range(len(models) - 1)
        ]
    else:
        assert len(transformation) == len(models) - 1, "The length of transformation should be len(models) - 1."
    if save_models_path is not None:
        Path(save_models_path).mkdir(parents=True, exist_ok=True)
    align_models = []
    cur_model = _read_h5ad(os.path.join(models_path, models[0])) if models_path is not None else models[0]
    cur_model.obsm[key_added] = cur_model.obsm[spatial_key].copy()
    if save_models_path is not None:
        cur_model.write(os.path.join(save_models_path, models[0]))
    elif models_path is not None:
        align_models.append(cur_model)
    for i, (cur_R, cur_t) in enumerate(compose_transformations(transformation)):
        cur_model = _read_h5ad(os.path.join(models_path, models[i + 1])) if models_path is not None else models[i + 1]
        cur_model.obsm[key_added] = cur_model.obsm[spatial_key].copy() @ cur_R.T + cur_t
        if save_models_path is not None:
            cur_model.write(os.path.join(save_models_path, models[i + 1]))
        elif models_path is not None:
            align_models.append(cur_model)
    return align_models if models_path is not None else models


def morpho_align_ref(
    models: List,
    models_ref: Optional[List] = None,
    n_sampling: Optional[int] = 2000,
    sampling_method: str = "random",
    rep_layer: Union[str, List[str]] = "X",
    rep_field: Union[str, List[str]] = "layer",
    genes: Optional[Union[list, np.ndarray]] = None,
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    iter_key_added: Optional[str] = "iter_spatial",
    vecfld_key_added: Optional[str] = "VecFld_morpho",
    mode: str = "SN-S",
    dissimilarity: Union[str, List[str]] = "kl",
    max_iter: int = 200,
    dtype: str = "float32",
    device: str = "cpu",
    verbose: bool = True,
    **kwargs,
):
    """Align down-sampled reference models, then carry the learned field to the full models with ``BA_transform``
    (morpho_alignment.py:318-454). Down-sampling: the reference delegates to third-party ``dynamo.tools.sampling``
    (absent); only ``sampling_method="random"`` is provided here."""
    if models_ref is None:
        if sampling_method != "random":
            raise NotImplementedError("only sampling_method='random' is available (dynamo's trn/kmeans samplers are third-party)")
        models_ref = []
        for m in models:
            n = m.shape[0]
            idx = np.sort(np.random.choice(n, min(n_sampling, n), replace=False))
            models_ref.append(m[idx].copy())
    pis, pis_ref = [], []
    align_models = [m.copy() for m in models]
    align_models_ref = [m.copy() for m in models_ref]
    for group in (align_models, align_models_ref):
        for m in group:
            m.obsm[key_added] = m.obsm[spatial_key].copy()
            m.obsm[f"{key_added}_rigid"] = m.obsm[spatial_key].copy()
            m.obsm[f"{key_added}_nonrigid"] = m.obsm[spatial_key].copy()
    for i in range(len(align_models) - 1):
        modelA_ref, modelB_ref = align_models_ref[i], align_models_ref[i + 1]
        morpho_model = Morpho_pairwise(
            sampleA=modelB_ref, sampleB=modelA_ref, rep_layer=rep_layer, rep_field=rep_field, dissimilarity=dissimilarity,
            genes=genes, spatial_key=key_added, key_added=key_added, iter_key_added=iter_key_added,
            vecfld_key_added=vecfld_key_added, max_iter=max_iter, dtype=dtype, device=device, verbose=verbose, **kwargs,
        )
        P = morpho_model.run()
        modelB_ref.obsm[f"{key_added}_rigid"] = morpho_model.optimal_RnA.copy()
        modelB_ref.obsm[f"{key_added}_nonrigid"] = morpho_model.XAHat.copy()
        modelB_ref.obsm[key_added] = modelB_ref.obsm[f"{key_added}_rigid" if mode == "SN-S" else f"{key_added}_nonrigid"]
        pis_ref.append(P)
        modelB = align_models[i + 1]
        if iter_key_added is not None:
            modelB_ref.uns[iter_key_added] = morpho_model.iter_added
            modelB.uns[iter_key_added] = morpho_model.iter_added
        if vecfld_key_added is not None:
            modelB_ref.uns[vecfld_key_added] = morpho_model.vecfld
            modelB.uns[vecfld_key_added] = morpho_model.vecfld
        modelB.obsm[f"{key_added}_nonrigid"], _, modelB.obsm[f"{key_added}_rigid"] = BA_transform(
            vecfld=morpho_model.vecfld, quary_points=modelB.obsm[key_added], device=device, dtype=dtype
        )
        modelB.obsm[key_added] = modelB.obsm[f"{key_added}_rigid" if mode == "SN-S" else f"{key_added}_nonrigid"]
        pis.append(P)
    return align_models, align_models_ref, pis, pis_ref


def remove_all_files_in_directory(directory_path):
    if os.path.exists(directory_path):
        for name in os.listdir(directory_path):
            fp = os.path.join(directory_path, name)
            if os.path.isfile(fp) or os.path.islink(fp):
                os.unlink(fp)
            elif os.path.isdir(fp):
                shutil.rmtree(fp)
