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


def align_chain_pipelined(
    get_slice: Callable[[int], object],
    n_slices: int,
    spatial_key: str = "spatial",
    key_added: str = "align_spatial",
    device=None,
    stats: Optional[dict] = None,
    **pairwise_kwargs,
):
    """Chain alignment of ``n_slices`` serial sections on this rank's GPU, software-pipelined (BASELINE configs[2]).

    Same result as ``morpho_align_chain_sharded`` (independent pairs on raw coordinates, morpho_alignment.py:181-217; ONE
    all-gather of the per-pair 2-D similarities; serial prefix composition, :300-303) but a rank that owns several pairs
    overlaps them: the EM of pair p is only *enqueued* on the main stream (CUDA-graph replays, no host waits), and while it
    runs the host prepares pair p + world on a second stream — constructor, coarse initialisation, the staged host-to-device
    copy of its two expression matrices and its cost matrix — so that the next EM starts as soon as the current one ends.

    ``get_slice(k)`` returns slice k (AnnData-like, host arrays); only the slices of this rank's pairs are requested, and
    only they receive ``obsm[key_added]``. Returns ``(placed, transformations)`` with ``placed`` = {slice index: slice}.
    ``stats`` (optional dict) receives per-pair timings.
    """
    import time

    from .morpho_class import Morpho_pairwise, resolve_device
    from .utils import solve_RT_by_correspondence

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = resolve_device(device)
    n_pairs = n_slices - 1
    mine = shard_pairs(n_pairs, rank, world)
    pairwise_kwargs.setdefault("materialize_P", False)
    cache = {}

    def slice_(k):
        if k not in cache:
            cache[k] = get_slice(k)
        return cache[k]

    def make(p):  # fixed = slice p, moving = slice p + 1
        s = Morpho_pairwise(sampleA=slice_(p + 1), sampleB=slice_(p), spatial_key=spatial_key, device=dev, **pairwise_kwargs)
        s.prepare()
        return s

    local = {}
    t_pairs = []
    from .. import _capi

    launches0, replayed = _capi.load_library().spb_launch_count(), 0
    with torch.cuda.device(dev):
        main, side = torch.cuda.current_stream(), torch.cuda.Stream()
        nxt = make(mine[0]) if mine else None
        for i, p in enumerate(mine):
            t0 = time.perf_counter()
            cur = nxt
            cur.run_em()  # enqueued only: the host is free while the device iterates
            nxt = None
            if i + 1 < len(mine):
                with torch.cuda.stream(side):
                    nxt = make(mine[i + 1])
            cur._finish()  # device -> host of the pair's results (waits for its EM)
            R, t = solve_RT_by_correspondence(cur.optimal_RnA[:, :2], np.asarray(slice_(p + 1).obsm[spatial_key])[:, :2])
            local[p] = {"Rotation": R, "Translation": t}
            replayed += getattr(cur, "graph_replayed_launches", 0)
            del cur
            main.wait_stream(side)
            t_pairs.append(time.perf_counter() - t0)
    transformation = gather_transformations(local, n_pairs, device=dev if dist.is_initialized() and dist.get_backend() == "nccl" else None)
    placed = {}
    composed = [(np.diag((1.0, 1.0)), np.zeros((2,)))] + compose_transformations(transformation)
    for k in sorted(cache):
        R, t = composed[k]
        sl = cache[k]
        raw = np.asarray(sl.obsm[spatial_key]).copy()
        sl.obsm[key_added] = raw if k == 0 else raw[:, :2] @ R.T + t
        placed[k] = sl
    if stats is not None:
        stats.update(pairs=list(mine), seconds_per_pair=t_pairs,
                     kernel_launches=int(_capi.load_library().spb_launch_count() - launches0 + replayed))
    return placed, transformation


def column_block(n_cols: int, rank: int, world: int):
    """Fixed cells [begin, end) of rank ``rank`` when one pair's columns are split over ``world`` GPUs."""
    return (n_cols * rank) // world, (n_cols * (rank + 1)) // world


_HOST_INIT_FIELDS = ("coordsA", "init_R", "init_t", "inlier_A", "inlier_B", "inlier_P", "sigma2", "_sigma2_init",
                     "probability_parameters", "samples_s", "outlier_s", "sigma2_variance_decress", "batch_perm")


def morpho_align_pair_sharded(fixed, moving, mode: str = "auto", device=None, **pairwise_kwargs):
    """ONE slice pair over all ranks of the process group ("single huge pair across GPUs").

    Every rank holds the whole moving slice (rows of P) and a block of the fixed slice's cells (columns of P): its block of
    the expression-probability matrix (N_A x N_B / world), sweep 1 and the column constants are local, and the only exchange is
    the sum of the per-row statistics of sweep 2 — [K_NA_spatial, K_NA_sigma2, sum P d, K_NA, P @ XB] = 7 fp64 per moving cell,
    5.6 MB at 100k cells — once per iteration. ``mode="p2p"`` (default when symmetric memory is available): that sum is done
    INSIDE the row-finalize kernel by reading the peers' partial vectors over NVLink in rank order (bit-identical replicas, no
    separate collective); ``mode="nccl"``: ``all_reduce`` + a finishing kernel (the baseline). The M-step runs replicated.

    Rank 0's host initialisation (coarse rigid alignment, sigma2 / beta2 guesses) is broadcast so the replicas start from the
    same bits. Returns the solver (same result attributes as ``Morpho_pairwise``; identical on every rank)."""
    from .morpho_class import Morpho_pairwise

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    kw = dict(pairwise_kwargs)
    kw.setdefault("materialize_P", False)
    kw["SVI_mode"] = False
    solver = Morpho_pairwise(sampleA=moving, sampleB=fixed, device=device, column_shard=(rank, world, mode), **kw)
    solver.prepare_host()
    if world > 1:
        box = [{k: getattr(solver, k) for k in _HOST_INIT_FIELDS if hasattr(solver, k)} if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        for k, v in box[0].items():
            setattr(solver, k, v)
    solver.prepare_device()
    return solver
