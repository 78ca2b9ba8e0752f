"""Multi-GPU chain alignment: independent slice pairs shard one per GPU, ONE all-gather of the per-pair rigid
transforms, then the serial prefix composition of the reference's ``morpho_align_apply_transformation``
(spateo/alignment/morpho_alignment.py:181-217 for the independent pairs, :300-303 for the composition).

One process per GPU (``torchrun``); the data path has no collective — the pairs are independent problems — and the only
exchange is ``[R(2x2) | t(2)]`` per pair (48 bytes), which is what BASELINE.json calls the global rigid-consensus step.
"""

from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .morpho_alignment import compose_transformations, pair_transformation


def shard_pairs(n_pairs: int, rank: int, world: int) -> List[int]:
    """Pair p (fixed = slice p, moving = slice p+1) runs on rank p mod world."""
    return [p for p in range(n_pairs) if p % world == rank]


def gather_transformations(local: dict, n_pairs: int, device=None) -> List[dict]:
    """``local`` maps pair index -> {"Rotation", "Translation"}; returns the full ordered list on every rank using a
    single ``all_gather`` of fixed-size per-rank slabs (NCCL on GPUs, gloo on CPU)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    per_rank = (n_pairs + world - 1) // world
    if dist.is_initialized() and dist.get_backend() == "nccl":
        from .morpho_class import resolve_device  # GPU-index strings ("0") / None -> torch.device("cuda", i)

        device = resolve_device(device)
    slab = torch.zeros((per_rank, 8), dtype=torch.float64, device=device)  # [pair index + 1, R00 R01 R10 R11, t0 t1, pad]
    for slot, p in enumerate(shard_pairs(n_pairs, rank, world)):
        tr = local[p]
        slab[slot, 0] = p + 1
        slab[slot, 1:5] = torch.as_tensor(np.asarray(tr["Rotation"], dtype=np.float64).reshape(-1)[:4])
        slab[slot, 5:7] = torch.as_tensor(np.asarray(tr["Translation"], dtype=np.float64).reshape(-1)[:2])
    if world > 1:
        out = torch.zeros((world, per_rank, 8), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(out.view(world * per_rank, 8), slab)
    else:
        out = slab[None]
    rows = out.reshape(-1, 8).cpu().numpy()
    result = [None] * n_pairs
    for r in rows:
        if r[0] > 0:
            result[int(round(r[0])) - 1] = {"Rotation": r[1:5].reshape(2, 2).copy(), "Translation": r[5:7].copy()}
    assert all(t is not None for t in result), "a pair transformation is missing after the all-gather"
    return result


def morpho_align_chain_sharded(
    models: List,
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    pair_fn: Optional[Callable] = None,
    device=None,
    **pairwise_kwargs,
):
    """Sharded equivalent of ``morpho_align_transformation`` + ``morpho_align_apply_transformation`` (rigid, 2-D, raw
    coordinates — NOT bitwise ``morpho_align``, whose pairs are serially dependent).

    Every rank receives the whole list of slices (or at least the ones it touches), aligns its pairs, joins the single
    all-gather, composes the chain and writes ``obsm[key_added]`` of every slice it holds. Returns
    ``(models, transformations)``.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_pairs = len(models) - 1
    pair_fn = pair_transformation if pair_fn is None else pair_fn
    local = {}
    for p in shard_pairs(n_pairs, rank, world):
        kw = dict(pairwise_kwargs)
        if device is not None and pair_fn is pair_transformation:
            kw.setdefault("device", device)
        local[p] = pair_fn(models[p], models[p + 1], spatial_key=spatial_key, **kw)
    transformation = gather_transformations(local, n_pairs, device=device if dist.is_initialized() and dist.get_backend() == "nccl" else None)
    models[0].obsm[key_added] = np.asarray(models[0].obsm[spatial_key]).copy()
    for i, (R, t) in enumerate(compose_transformations(transformation)):
        m = models[i + 1]
        m.obsm[key_added] = np.asarray(m.obsm[spatial_key]).copy()[:, :2] @ R.T + t
    return models, transformation
