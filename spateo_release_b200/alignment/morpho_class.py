"""``Morpho_pairwise`` — drop-in for ``spateo.alignment.methods.morpho_class.Morpho_pairwise`` (morpho_class.py:54)
whose EM runs as hand-written sm_100a CUDA kernels behind the C ABI in ``include/spateo_b200.h``.

Host side (this file): validation, gene intersection, dense extraction, coordinate normalisation, inducing-point choice,
coarse rigid initialisation bookkeeping and output wrapping — same names, argument meaning, RNG call order
(SURVEY.md Appendix D), result attributes and exceptions as the reference. Device side: expression-cost matrix,
fused two-sweep E-step (P never materialised), gamma/alpha, non-rigid solve, rigid Procrustes, sigma2; all scalar state
stays on the device, there is no host synchronisation inside the iteration loop and no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Union

import numpy as np
import torch

from .. import _capi
from .._capi import SpbEmParams, SpbScalars, check, ptr
from . import utils as U

_METRIC_CODE = {
    "kl": _capi.CONST["SPB_METRIC_KL"],
    "euc": _capi.CONST["SPB_METRIC_EUC"],
    "euclidean": _capi.CONST["SPB_METRIC_EUC"],
    "square_euc": _capi.CONST["SPB_METRIC_SQRT_EUC"],
    "square_euclidean": _capi.CONST["SPB_METRIC_SQRT_EUC"],
    "cos": _capi.CONST["SPB_METRIC_COS"],
    "cosine": _capi.CONST["SPB_METRIC_COS"],
    "sym_kl": _capi.CONST["SPB_METRIC_SYMKL"],
}
_PROB_CODE = {
    "gauss": _capi.CONST["SPB_PROB_GAUSS"],
    "gaussian": _capi.CONST["SPB_PROB_GAUSS"],
    "cos": _capi.CONST["SPB_PROB_COS"],
    "cosine": _capi.CONST["SPB_PROB_COS"],
    "prob": _capi.CONST["SPB_PROB_PROB"],
}


# bytes moved between host and device by the big transfers of an alignment (expression matrices, coordinate / result
# arrays); bench.py reads and resets it around the public call for the end-to-end line
TRANSFER_BYTES = {"h2d": 0, "d2h": 0}


def _count_h2d(t) -> None:
    TRANSFER_BYTES["h2d"] += int(t.numel()) * int(t.element_size())


def _count_d2h(t) -> None:
    TRANSFER_BYTES["d2h"] += int(t.numel()) * int(t.element_size())


_STAGE_FLOATS = 16 << 20  # two reusable 64 MB pinned staging buffers
_stage = {}
_stage_lock = __import__("threading").Lock()


def staged_to_device(host: np.ndarray, dev: torch.device) -> torch.Tensor:
    """Pageable host array -> device through two reusable pinned staging buffers (chunk k is copied into pinned memory
    while chunk k-1 is on the wire): 36 GB/s measured for an 800 MB expression matrix against 11 GB/s for a pageable
    ``.to()`` and a 0.6 s first-use cost for ``pin_memory()`` (profiles/h2d_micro.py). Counts the bytes in
    ``TRANSFER_BYTES``."""
    t = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32))
    _count_h2d(t)
    if t.is_pinned() or t.numel() < (2 << 20):
        return t.to(dev, non_blocking=t.is_pinned())
    events = [torch.cuda.Event(), torch.cuda.Event()]
    out = torch.empty(t.shape, dtype=torch.float32, device=dev)
    src, dst = t.view(-1), out.view(-1)
    n = src.numel()
    with _stage_lock, torch.cuda.device(dev):
        if "bufs" not in _stage:
            _stage["bufs"] = [torch.empty((_STAGE_FLOATS,), dtype=torch.float32).pin_memory() for _ in range(2)]
        bufs = _stage["bufs"]
        for k, o in enumerate(range(0, n, _STAGE_FLOATS)):
            b = k & 1
            m = min(_STAGE_FLOATS, n - o)
            if k >= 2:
                events[b].synchronize()
            bufs[b][:m].copy_(src[o:o + m])
            dst[o:o + m].copy_(bufs[b][:m], non_blocking=True)
            events[b].record()
        for e in events[: min(2, (n + _STAGE_FLOATS - 1) // _STAGE_FLOATS)]:
            e.synchronize()  # the staging buffers are reused by the next call
    return out


def _round_up(x: int, m: int) -> int:
    return ((x + m - 1) // m) * m


class _nvtx:
    """NVTX range (``SPB_NVTX=1``) around the phases of an alignment, for nsys / ncu --nvtx timelines."""

    enabled = os.environ.get("SPB_NVTX", "0") == "1"

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if self.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if self.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def morton_order(coords: np.ndarray) -> np.ndarray:
    """Row permutation that sorts points along a Z-order curve (isotropic quantisation: 16 bits/axis in 2-D, 10 in 3-D)."""
    c = np.asarray(coords, dtype=np.float64)
    n, D = c.shape
    bits = 16 if D == 2 else 10
    lo = c.min(axis=0)
    ext = max(float((c.max(axis=0) - lo).max()), 1e-30)
    q = np.minimum(((c - lo) / ext * (2**bits - 1)).astype(np.uint64), np.uint64(2**bits - 1))
    code = np.zeros(n, dtype=np.uint64)
    for b in range(bits):
        for d in range(D):
            code |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * D + d)
    return np.argsort(code, kind="stable")


def resolve_device(device) -> torch.device:
    """Reference semantics: "cpu" or a GPU index string (utils.py:35-66). Here every value maps to a CUDA device —
    there is no CPU path; ``CUDA_VISIBLE_DEVICES`` is NOT mutated (the reference does, utils.py:51)."""
    _capi.require_cuda()
    if isinstance(device, torch.device):
        return device
    if device is None or device == "cpu" or device == "cuda":
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        return torch.device("cuda", device)
    s = str(device)
    if s.startswith("cuda:"):
        return torch.device(s)
    return torch.device("cuda", int(s))


class GeneCostBuilder:
    """Device pipeline for ``calc_distance`` + ``calc_probability`` of one representation layer (utils.py:866-985).

    Two contraction back-ends behind the same call: ``tensor`` (default) = tcgen05 / TMEM / TMA kernel with a 3xTF32
    error-compensated split (fp32-accurate), ``simt`` = packed-FFMA2 register-tiled kernel with two-level accumulation.
    """

    def __init__(self, lib, dev, backend: Optional[str] = None):
        import os

        self.lib, self.dev = lib, dev
        self.backend = backend or os.environ.get("SPB_GENE_COST", "tensor")

    def prepare(self, X: torch.Tensor, metric: str, fixed: bool, centre: Optional[torch.Tensor] = None):
        """Row pre-pass. Returns (operand [n, Gp] fp32 zero-padded to 32 features, rowterm [n] or None).

        KL, fixed side: ``centre`` (the mean normalised moving profile, length G) centres every log-row so the contraction
        stays near zero (see kl_prepare_rows_kernel); the centring term comes back as the row term."""
        n, G = X.shape
        Gp = _round_up(G, 32)
        st = _capi.current_stream_ptr()
        if metric == "kl":
            out = torch.empty((n, Gp), dtype=torch.float32, device=self.dev)
            want_rt = (not fixed) or (centre is not None)
            rt = torch.empty((n,), dtype=torch.float32, device=self.dev) if want_rt else None
            check(self.lib.spb_kl_prepare_rows(ptr(X), n, G, X.stride(0), ptr(out), Gp, ptr(rt), 1 if fixed else 0,
                                               ptr(centre), st), "spb_kl_prepare_rows")
            return out, rt
        if metric in ("cos", "cosine"):
            out = torch.empty((n, Gp), dtype=torch.float32, device=self.dev)
            check(self.lib.spb_rows_normalize(ptr(X), n, G, X.stride(0), ptr(out), Gp, st), "spb_rows_normalize")
            return out, None
        # euclidean family: raw values (padded) + squared norms
        if Gp == G and X.is_contiguous():
            out = X
        else:
            out = torch.zeros((n, Gp), dtype=torch.float32, device=self.dev)
            out[:, :G] = X
        rt = torch.empty((n,), dtype=torch.float32, device=self.dev)
        check(self.lib.spb_rows_sqnorm(ptr(out), n, G, Gp, ptr(rt), st), "spb_rows_sqnorm")
        return out, rt

    def prepare_pair(self, A: torch.Tensor, B: torch.Tensor, metric: str):
        """Operands + row terms of (moving A, fixed B) for ``metric``. Returns (opA, rtA, opB, rtB, G_effective).

        ``sym_kl`` = (KL(a||b) + KL(b||a)) / 2 (utils.py:922-932) is ONE contraction over 2G features:
        [Xn | log X - d_i] . [log Y - c_j | Yn], with both log blocks centred (c_j by the mean moving profile, d_i by the
        mean fixed profile) and e = ((sum Xn log X - d_i) + (sum Yn log Y - c_j) - dot) / 2."""
        G = A.shape[1]
        if metric != "sym_kl":
            opA, rtA = self.prepare(A, metric, fixed=False)
            opB, rtB = self.prepare(B, metric, fixed=True, centre=self.centre_of(opA, G) if metric == "kl" else None)
            return opA, rtA, opB, rtB, G
        Xn, ta = self.prepare(A, "kl", fixed=False)              # Xn, sum Xn (log Xn + log G)
        Yn, tb = self.prepare(B, "kl", fixed=False)
        LY, cj = self.prepare(B, "kl", fixed=True, centre=self.centre_of(Xn, G))   # log Y + log G - c_j
        LX, di = self.prepare(A, "kl", fixed=True, centre=self.centre_of(Yn, G))   # log X + log G - d_i
        Gp = Xn.shape[1]
        opA = torch.cat([Xn, LX], dim=1).contiguous()
        opB = torch.cat([LY, Yn], dim=1).contiguous()
        return opA, (ta - di).contiguous(), opB, (tb - cj).contiguous(), 2 * Gp

    @staticmethod
    def centre_of(opA: torch.Tensor, G: int) -> torch.Tensor:
        """Mean normalised moving profile (fp32, length G) used to centre the fixed side of the KL contraction."""
        return opA[:, :G].mean(dim=0, dtype=torch.float64).float().contiguous()

    def _split(self, op: torch.Tensor):
        hi, lo = torch.empty_like(op), torch.empty_like(op)
        check(self.lib.spb_split_tf32(ptr(op), ptr(hi), ptr(lo), op.numel(), _capi.current_stream_ptr()), "spb_split_tf32")
        return hi, lo

    def cost(self, opA, rtA, opB, rtB, NA, NB, G, metric, prob_type, prob_param, accumulate, GT, ldx, backend=None):
        backend = backend or self.backend
        pp = float(prob_param) if prob_param is not None else 1.0
        if backend == "tensor":
            ahi, alo = self._split(opA)
            bhi, blo = self._split(opB)
            check(
                self.lib.spb_gene_cost_tc(
                    ptr(ahi), ptr(alo), opA.stride(0), ptr(rtA), ptr(bhi), ptr(blo), opB.stride(0), ptr(rtB), NA, NB, G,
                    _METRIC_CODE[metric], _PROB_CODE[prob_type], pp, 1 if accumulate else 0, ptr(GT), ldx,
                    _capi.current_stream_ptr(),
                ),
                "spb_gene_cost_tc",
            )
            self._keep = (ahi, alo, bhi, blo)  # stay alive until the stream has consumed them
            return
        check(
            self.lib.spb_gene_cost(
                ptr(opA), opA.stride(0), ptr(rtA), ptr(opB), opB.stride(0), ptr(rtB), NA, NB, G, _METRIC_CODE[metric],
                _PROB_CODE[prob_type], pp, 1 if accumulate else 0, ptr(GT), ldx, _capi.current_stream_ptr(),
            ),
            "spb_gene_cost",
        )


class Morpho_pairwise:
    """Align a moving slice ``sampleA`` onto a fixed slice ``sampleB`` (same constructor as morpho_class.py:110-167).

    Extra keywords (not in the reference): ``materialize_P`` — when False ``run()`` skips building the dense
    N_A x N_B posterior (40 GB at 100k x 100k) and returns None; every other output is unaffected.
    ``compute_mapping`` — ``self.mapping`` (an ``ArgmaxPi``) receives the row / column maxima of the final posterior from a
    fused kernel, for ``get_optimal_mapping_relationship`` / ``mapping_aligned_coords`` without a dense P.
    ``spatial_sort`` / ``cull_zero_tiles`` — the moving cells are processed in Morton order so that each row block (SPB_ROW_TILE = 512 cells) is
    spatially compact, and (row block, fixed cell) tiles whose every pair underflows to exactly 0 in fp32 are skipped;
    results are bit-identical to the dense sweep (all outputs are returned in the caller's row order).
    ``column_shard`` — set by ``morpho_align_pair_sharded``: one pair's fixed cells split over several GPUs.
    Accepted but without effect (memory work-arounds whose results are identical): ``use_chunk``, ``chunk_capacity``,
    ``pre_compute_dist``. ``sparse_calculation_mode`` keeps the top ``sparse_top_k`` posterior entries of every column by an
    exact on-device radix select (P comes back as ``scipy.sparse.coo_matrix``). Not implemented (NotImplementedError):
    ``kernel_type="geodist"``.
    """

    def __init__(
        self,
        sampleA,
        sampleB,
        rep_layer: Union[str, List[str]] = "X",
        rep_field: Union[str, List[str]] = "layer",
        genes=None,
        spatial_key: str = "spatial",
        key_added: str = "align_spatial",
        iter_key_added: Optional[str] = None,
        save_concrete_iter: bool = False,
        vecfld_key_added: Optional[str] = None,
        dissimilarity: Union[str, List[str]] = "kl",
        probability_type: Union[str, List[str]] = "gauss",
        probability_parameters=None,
        label_transfer_dict=None,
        use_hvg: bool = True,
        nn_init: bool = True,
        init_transform: bool = True,
        allow_flip: bool = False,
        init_layer: str = "X",
        init_field: str = "layer",
        nn_init_top_K: int = 10,
        nn_init_weight: float = 1.0,
        max_iter: int = 200,
        nonrigid_start_iter: int = 80,
        SVI_mode: bool = True,
        batch_size: Optional[int] = None,
        pre_compute_dist: bool = True,
        sparse_calculation_mode: bool = False,
        sparse_top_k: int = 1024,
        lambdaVF: Union[int, float] = 1e2,
        beta: Union[int, float] = 0.01,
        K: Union[int, float] = 15,
        kernel_type: str = "euc",
        graph=None,
        graph_knn: int = 10,
        sigma2_init_scale: Optional[Union[int, float]] = 0.1,
        sigma2_end: Optional[Union[int, float]] = None,
        gamma_a: float = 1.0,
        gamma_b: float = 1.0,
        kappa: Union[float, np.ndarray] = 1.0,
        partial_robust_level: float = 10,
        normalize_c: bool = True,
        normalize_g: bool = False,
        separate_mean: bool = True,
        separate_scale: bool = False,
        dtype: str = "float32",
        device: str = "cpu",
        verbose: bool = True,
        guidance_pair=None,
        guidance_effect=False,
        guidance_weight: float = 1.0,
        use_chunk: bool = False,
        chunk_capacity: float = 1.0,
        return_mapping: bool = False,
        update_R: bool = True,
        materialize_P: bool = True,
        compute_mapping: bool = False,
        spatial_sort: bool = True,
        cull_zero_tiles: bool = True,
        column_shard=None,
    ) -> None:
        self.verbose = verbose
        self.sampleA, self.sampleB = sampleA, sampleB
        self.rep_layer, self.rep_field, self.genes = rep_layer, rep_field, genes
        self.spatial_key, self.key_added = spatial_key, key_added
        self.iter_key_added, self.save_concrete_iter = iter_key_added, save_concrete_iter
        self.vecfld_key_added = vecfld_key_added
        self.dissimilarity, self.probability_type = dissimilarity, probability_type
        self.probability_parameters = probability_parameters
        self.label_transfer_dict = label_transfer_dict
        self.use_hvg, self.nn_init, self.init_transform = use_hvg, nn_init, init_transform
        self.nn_init_top_K, self.max_iter, self.allow_flip = nn_init_top_K, max_iter, allow_flip
        self.init_layer, self.init_field = init_layer, init_field
        self.SVI_mode, self.batch_size, self.pre_compute_dist = SVI_mode, batch_size, pre_compute_dist
        self.sparse_calculation_mode, self.sparse_top_k = sparse_calculation_mode, sparse_top_k
        self.beta, self.lambdaVF, self.K = beta, lambdaVF, int(K)
        self.kernel_type, self.kernel_bandwidth = kernel_type, beta
        self.graph, self.graph_knn = graph, graph_knn
        self.sigma2_init_scale, self.sigma2_end = sigma2_init_scale, sigma2_end
        self.partial_robust_level = partial_robust_level
        self.normalize_c, self.normalize_g = normalize_c, normalize_g
        self.separate_mean, self.separate_scale = separate_mean, separate_scale
        self.dtype, self.device = dtype, device
        self.guidance_pair, self.guidance_effect, self.guidance_weight = guidance_pair, guidance_effect, guidance_weight
        self.use_chunk, self.chunk_capacity = use_chunk, chunk_capacity
        self.nn_init_weight = nn_init_weight
        self.gamma_a, self.gamma_b, self.kappa = gamma_a, gamma_b, kappa
        self.nonrigid_start_iter = nonrigid_start_iter
        self.return_mapping, self.update_R = return_mapping, update_R
        self.materialize_P = materialize_P
        self.compute_mapping = compute_mapping
        self.spatial_sort, self.cull_zero_tiles = spatial_sort, cull_zero_tiles
        self.use_cuda_graph = os.environ.get("SPB_CUDA_GRAPH", "1") != "0"
        # iterations per captured graph: light iterations (SVI batches, small pairs) are bound by the host's graph launches
        # on a slow host, so several identical iterations ride in one graph; 0 = choose from the pairs per iteration
        self.graph_unroll = int(os.environ.get("SPB_GRAPH_UNROLL", "0"))
        # column-sharded pair: (rank, world, mode) — this process holds the fixed cells [NB * rank / world, NB * (rank + 1) /
        # world) of ONE pair; see alignment/distributed.py:morpho_align_pair_sharded
        self.column_shard = column_shard

        self._np_dtype = np.float32 if dtype == "float32" else np.float64
        self._check()
        self._lib = _capi.load_library()
        self._dev = resolve_device(device)
        with torch.cuda.device(self._dev):
            self._align_preprocess()
            self._construct_kernel()

    # ------------------------------------------------------------------------------------------------------------------
    # validation (morpho_class.py:316-440)
    # ------------------------------------------------------------------------------------------------------------------
    def _check(self):
        if self.rep_layer is None:
            raise ValueError(
                "No representation input is detected, which may not produce meaningful result. Please check the rep_layer and rep_field."
            )
        if self.rep_field is None:
            self.rep_field = "layer"
        if isinstance(self.rep_layer, str):
            self.rep_layer = [self.rep_layer]
        if isinstance(self.rep_field, str):
            self.rep_field = [self.rep_field] * len(self.rep_layer)
        if not U.check_rep_layer([self.sampleA, self.sampleB], self.rep_layer, self.rep_field):
            raise ValueError("The specified representation is not found in the attribute of the AnnData objects.")
        self.obs_key = U.check_obs(self.rep_layer, self.rep_field)
        if self.spatial_key not in self.sampleA.obsm:
            raise KeyError(f"Spatial key '{self.spatial_key}' not found in sampleA AnnData object.")
        if self.spatial_key not in self.sampleB.obsm:
            raise KeyError(f"Spatial key '{self.spatial_key}' not found in sampleB AnnData object.")
        if self.obs_key is not None and self.label_transfer_dict is not None:
            catA = self.sampleA.obs[self.obs_key].cat.categories.tolist()
            catB = self.sampleB.obs[self.obs_key].cat.categories.tolist()
            U.check_label_transfer_dict(catA, catB, self.label_transfer_dict)
        if self.dissimilarity is None:
            self.dissimilarity = "kl"
        if isinstance(self.dissimilarity, str):
            self.dissimilarity = [self.dissimilarity] * len(self.rep_layer)
        valid = ["kl", "sym_kl", "euc", "euclidean", "square_euc", "square_euclidean", "cos", "cosine", "label"]
        self.dissimilarity = [d.lower() for d in self.dissimilarity]
        for d in self.dissimilarity:
            if d not in valid:
                raise ValueError(f"Invalid `metric` value: {d}. Available `metrics` are: " f"{', '.join(valid)}.")
        if self.probability_type is None:
            self.probability_type = "gauss"
        if isinstance(self.probability_type, str):
            self.probability_type = [self.probability_type] * len(self.rep_layer)
        validp = ["gauss", "gaussian", "cos", "cosine", "prob"]
        self.probability_type = [p.lower() for p in self.probability_type]
        for p in self.probability_type:
            if p not in validp:
                raise ValueError(f"Invalid `metric` value: {p}. Available `metrics` are: " f"{', '.join(validp)}.")
        for i, f in enumerate(self.rep_field):
            if f == "obs":
                self.dissimilarity[i] = "label"
                self.probability_type[i] = "prob"
        if self.probability_parameters is None:
            self.probability_parameters = [None] * len(self.rep_layer)
        elif not isinstance(self.probability_parameters, (list, tuple)):
            self.probability_parameters = [self.probability_parameters] * len(self.rep_layer)
        self.probability_parameters = list(self.probability_parameters)
        if self.nn_init:
            if not U.check_rep_layer([self.sampleA, self.sampleB], [self.init_layer], [self.init_field]):
                raise ValueError("The specified representation is not found in the attribute of the AnnData objects.")
        if self.guidance_effect:
            valid_g = ["nonrigid", "rigid", "both"]
            if self.guidance_effect not in valid_g:
                raise ValueError(
                    f"Invalid `guidance_effect` value: {self.guidance_effect}. Available `guidance_effect` values are: "
                    f"{', '.join(valid_g)}."
                )
        # ---- features of the reference that this round does not cover: fail loudly, never silently differ ----
        if self.sparse_calculation_mode:
            self.pre_compute_dist = False  # morpho_class.py:439-440 (no effect here: the cost matrix is always resident)
            if int(self.sparse_top_k) < 1:
                raise ValueError("sparse_top_k must be a positive integer.")
        if self.kernel_type not in ("euc", "geodist"):
            raise NotImplementedError(f"Kernel type '{self.kernel_type}' is not implemented.")
        if self.dtype != "float32":
            # the reference honours dtype="float64" end to end (morpho_class.py:165, utils.py:35-66); the device kernels of
            # this package compute in float32 (with fp64 reductions), so a float64 request is refused rather than served
            # with narrower arithmetic
            raise NotImplementedError(
                f"dtype={self.dtype!r} is not implemented in spateo_release_b200: the device path computes in float32 "
                "(fp64 reductions / solves); use dtype='float32'."
            )

    # ------------------------------------------------------------------------------------------------------------------
    # preprocessing (morpho_class.py:443-558)
    # ------------------------------------------------------------------------------------------------------------------
    def _align_preprocess(self):
        dt = self._np_dtype
        if self.use_hvg and ("highly_variable" in self.sampleA.var.columns) and ("highly_variable" in self.sampleB.var.columns):
            gl = [
                self.sampleA.var.index[self.sampleA.var.highly_variable],
                self.sampleB.var.index[self.sampleB.var.highly_variable],
            ]
        else:
            gl = [self.sampleA.var.index, self.sampleB.var.index]
        common = U.filter_common_genes(*gl, verbose=self.verbose)
        self.genes = common if self.genes is None else U.intersect_lsts(common, list(self.genes))

        self.exp_layers_A = [U.get_rep(self.sampleA, r, f, self.genes, dt) for r, f in zip(self.rep_layer, self.rep_field)]
        self.exp_layers_B = [U.get_rep(self.sampleB, r, f, self.genes, dt) for r, f in zip(self.rep_layer, self.rep_field)]
        if self.obs_key is not None:
            self.label_transfer = U.check_label_transfer(self.sampleA, self.sampleB, self.obs_key, self.label_transfer_dict)
        else:
            self.label_transfer = None

        self.coordsA = U.check_spatial_coords(self.sampleA, self.spatial_key).astype(dt)
        self.coordsB = U.check_spatial_coords(self.sampleB, self.spatial_key).astype(dt)
        assert self.coordsA.shape[1] == self.coordsB.shape[1], "Spatial coordinate dimensions are different, please check again."
        self.NA, self.NB, self.D = self.coordsA.shape[0], self.coordsB.shape[0], self.coordsA.shape[1]
        if self.normalize_c:
            self.coordsA, self.coordsB, self.normalize_scales, self.normalize_means = U.normalize_coords(
                self.coordsA, self.coordsB, self.separate_mean, self.separate_scale
            )
        if self.normalize_g:
            self._normalize_exps()
        # guidance pairs [X_BI on the fixed slice, X_AI on the moving slice] (morpho_class.py:551-587)
        if (self.guidance_pair is not None) and (self.guidance_effect != False) and (self.guidance_weight > 0):  # noqa: E712
            if not isinstance(self.guidance_pair, list) or len(self.guidance_pair) != 2:
                raise ValueError("guidance_pair must be a list with two elements: [X_BI, X_AI].")
            self.X_BI = np.asarray(self.guidance_pair[0]).astype(dt)
            self.X_AI = np.asarray(self.guidance_pair[1]).astype(dt)
            if self.normalize_c:
                self.X_AI = (self.X_AI - self.normalize_means[0]) / self.normalize_scales[0]
                self.X_BI = (self.X_BI - self.normalize_means[1]) / self.normalize_scales[1]
            self.guidance = True
        else:
            self.guidance = False

    def _normalize_exps(self):
        """morpho_class.py:657-680: shared RMS scale for 'layer' representations whose metric is not KL."""
        for i, (f, d) in enumerate(zip(self.rep_field, self.dissimilarity)):
            if f == "layer" and d != "kl":
                sc = 0.0
                for e in (self.exp_layers_A[i], self.exp_layers_B[i]):
                    sc += np.sqrt(np.sum(e.astype(np.float64) ** 2) / e.shape[0])
                sc /= 2
                self.exp_layers_A[i] = (self.exp_layers_A[i] / sc).astype(self._np_dtype)
                self.exp_layers_B[i] = (self.exp_layers_B[i] / sc).astype(self._np_dtype)

    # ------------------------------------------------------------------------------------------------------------------
    # inducing points + kernel (morpho_class.py:825-875)
    # ------------------------------------------------------------------------------------------------------------------
    def _construct_kernel(self):
        uniq, uniq_idx = np.unique(self.coordsA, return_index=True, axis=0)
        if uniq.shape[0] > self.K:
            pick = np.random.choice(uniq.shape[0], self.K, replace=False)
        else:
            pick = np.arange(uniq.shape[0])
        self.inducing_variables_idx = uniq_idx[pick]
        self.inducing_variables = self.coordsA[self.inducing_variables_idx, :]
        self.K = self.inducing_variables.shape[0]
        z = self.inducing_variables.astype(np.float64)
        d2 = ((z[:, None, :] - z[None, :, :]) ** 2).sum(-1)
        self.GammaSparse = np.exp(-self.kernel_bandwidth * d2).astype(np.float32)
        if self.guidance and self.guidance_effect in ("nonrigid", "both"):
            xa = self.X_AI.astype(np.float64)
            self.U_I = np.exp(-self.kernel_bandwidth * ((xa[:, None, :] - z[None, :, :]) ** 2).sum(-1))  # [N_I, K] fp64
        else:
            self.U_I = None
        # U^T on the device from the pre-initialisation coordinates (the reference builds U before the coarse init)
        self.ldx = _round_up(self.NA, _capi.ROW_TILE)
        dev = self._dev
        if self.kernel_type == "geodist":
            self._construct_geodesic_kernel()
            return
        x_soa = torch.zeros((3, self.ldx), dtype=torch.float32, device=dev)
        x_soa[: self.D, : self.NA] = torch.from_numpy(np.ascontiguousarray(self.coordsA.T, dtype=np.float32)).to(dev)
        zt = torch.zeros((self.K, 3), dtype=torch.float32, device=dev)
        zt[:, : self.D] = torch.from_numpy(self.inducing_variables.astype(np.float32)).to(dev)
        self._UT = torch.empty((self.K, self.ldx), dtype=torch.float32, device=dev)
        check(
            self._lib.spb_rbf_kernel_T(ptr(x_soa), self.NA, self.ldx, ptr(zt), self.K, float(self.kernel_bandwidth),
                                       ptr(self._UT), _capi.current_stream_ptr()),
            "spb_rbf_kernel_T",
        )

    def _construct_geodesic_kernel(self):
        """``kernel_type="geodist"`` (morpho_class.py:865-871, utils.py:1161-1217): U = exp(-beta d_g^2) with d_g the
        shortest-path distance on the k-nearest-neighbour graph of the moving cells, from every cell to the K inducing
        cells; unreachable pairs get d_g = 1e5 like the reference. One-off host work (scipy's Dijkstra on the sparse graph
        instead of the reference's dense N x N adjacency + networkx loop), the kernel matrix then lives on the device."""
        import scipy.sparse as sp
        from scipy.sparse.csgraph import dijkstra

        N = self.NA
        if self.graph is None:
            from sklearn.neighbors import kneighbors_graph

            adj = kneighbors_graph(self.coordsA, self.graph_knn, mode="distance", include_self=False)
        elif sp.issparse(self.graph):
            adj = sp.csr_matrix(self.graph)
        else:  # a networkx graph, as the reference accepts
            import networkx

            adj = networkx.to_scipy_sparse_array(self.graph, nodelist=list(range(N)), weight="weight", format="csr")
        adj = sp.csr_matrix(adj.maximum(adj.T))  # networkx.Graph is undirected: an edge exists if either end lists it
        dist = dijkstra(adj, directed=False, indices=np.asarray(self.inducing_variables_idx, dtype=np.int64))  # [K, N]
        dist = np.where(np.isfinite(dist), dist, 1e5)
        UT64 = np.exp(-float(self.kernel_bandwidth) * dist**2)  # [K, N] float64 like the reference's
        self.GammaSparse = np.ascontiguousarray(UT64[:, self.inducing_variables_idx].T, dtype=np.float32)
        self.U_I = None  # guidance points are not nodes of the graph (morpho_class.py:871)
        self._UT = torch.zeros((self.K, self.ldx), dtype=torch.float32, device=self._dev)
        self._UT[:, :N] = torch.from_numpy(np.ascontiguousarray(UT64, dtype=np.float32)).to(self._dev)

    @property
    def U(self) -> np.ndarray:
        """[N_A, K] kernel matrix as the reference exposes it."""
        u = self._UT[:, : self.NA].T.contiguous().cpu().numpy().astype(self._np_dtype)
        return self._unsorted(u) if getattr(self, "_UT_is_sorted", False) else u

    # ------------------------------------------------------------------------------------------------------------------
    # device-side expression distances for small helper problems (coarse init, beta^2 init)
    # ------------------------------------------------------------------------------------------------------------------
    def _raw_cost_T(self, XA_host, XB_host, metric):
        """E^T[j][i] = metric(A_i, B_j) as a device tensor [nB, ldx_s] (columns beyond nA are padding).
        Inputs: host arrays or device tensors (the device voxel means are passed straight through)."""
        dev = self._dev
        gc = GeneCostBuilder(self._lib, dev)

        def up(x):
            if torch.is_tensor(x):
                return x.to(device=dev, dtype=torch.float32).contiguous()
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            _count_h2d(t)
            return t.to(dev)

        A, B = up(XA_host), up(XB_host)
        opA, rtA, opB, rtB, G = gc.prepare_pair(A, B, metric)
        nA, nB = A.shape[0], B.shape[0]
        lds = _round_up(nA, 256)
        ET = torch.empty((nB, lds), dtype=torch.float32, device=dev)
        gc.cost(opA, rtA, opB, rtB, nA, nB, G, metric, "prob", None, False, ET, lds)
        return ET, nA

    # ------------------------------------------------------------------------------------------------------------------
    # coarse rigid alignment (morpho_class.py:898-1041): voxelise, mutual top-K by expression, robust Procrustes
    # ------------------------------------------------------------------------------------------------------------------
    def _coarse_rigid_alignment(self, n_sampling: int = 20000):
        top_K = self.nn_init_top_K
        ia = np.random.choice(self.NA, n_sampling, replace=False) if self.NA > n_sampling else np.arange(self.NA)
        ib = np.random.choice(self.NB, n_sampling, replace=False) if self.NB > n_sampling else np.arange(self.NB)
        cA, cB = self.coordsA[ia, :], self.coordsB[ib, :]
        N, M, D = cA.shape[0], cB.shape[0], cA.shape[1]
        import time as _time

        _t = _time.perf_counter()
        XA = self._device_rows(self.init_layer, self.init_field, "A", ia)
        XB = self._device_rows(self.init_layer, self.init_field, "B", ib)
        if XA is None or XB is None:  # an initialisation layer that is not part of the alignment: host extraction
            XA = U.get_rep(self.sampleA[ia], self.init_layer, self.init_field, self.genes, self._np_dtype)
            XB = U.get_rep(self.sampleB[ib], self.init_layer, self.init_field, self.genes, self._np_dtype)
        # the first use of the representations uploads them (staged H2D of both expression matrices) — the one upload of a run
        self._timing["coarse.upload_and_gather_s"] = _time.perf_counter() - _t
        _t = _time.perf_counter()
        cA, XA = self._voxel_data_device(cA, XA, voxel_num=max(min(int(N / 20), 1000), 100))
        cB, XB = self._voxel_data_device(cB, XB, voxel_num=max(min(int(M / 20), 1000), 100))
        self._timing["coarse.voxel_data_s"] = _time.perf_counter() - _t
        _t = _time.perf_counter()
        metric = "kl" if self.init_field == "layer" else "euc"
        ET, nA = self._raw_cost_T(XA, XB, metric)  # [nB, lds]: ET[b, a] = dist(voxel a of A, voxel b of B)
        ET = ET[:, :nA]
        nB = ET.shape[0]
        while True:
            try:
                if top_K > nA - 1 or top_K > nB - 1:  # np.argpartition(kth=top_K) needs kth < size
                    raise ValueError(f"kth(={top_K}) out of bounds")
                # for every voxel b of B: the top_K voxels a of A (reference: argpartition over axis 0 of [nA, nB])
                d1, a_idx = torch.topk(ET, top_K, dim=1, largest=False)
                NN1 = np.stack([np.repeat(np.arange(nB), top_K), a_idx.reshape(-1).cpu().numpy()], axis=1)
                dist1 = d1.reshape(-1).cpu().numpy()
                # for every voxel a of A: the top_K voxels b of B
                d2, b_idx = torch.topk(ET, top_K, dim=0, largest=False)
                NN2 = np.stack([b_idx.T.reshape(-1).cpu().numpy(), np.repeat(np.arange(nA), top_K)], axis=1)
                dist2 = d2.T.reshape(-1).cpu().numpy()
                break
            except Exception as e:
                top_K -= 1
                if top_K == 0:
                    raise RuntimeError("Failed to perform coarse rigid alignment after reducing top_K.") from e
        NN = np.vstack((NN1, NN2))
        distance = np.r_[dist1, dist2].astype(np.float64)
        train_x, train_y = cA[NN[:, 1], :], cB[NN[:, 0], :]
        self._timing["coarse.expression_knn_s"] = _time.perf_counter() - _t
        _t = _time.perf_counter()
        P, R, t, sigma2, gamma = self._inlier_from_NN_device(train_x, train_y, distance)
        self._timing["coarse.inlier_from_NN_s"] = _time.perf_counter() - _t
        self._timing["coarse.n_pairs"] = int(train_x.shape[0])
        if self.allow_flip:
            Rf = np.eye(D)
            Rf[-1, -1] = -1
            P2, R2, t2, s2, g2 = self._inlier_from_NN_device(train_x @ Rf, train_y, distance)
            if g2 > gamma:
                P, R, t, sigma2 = P2, R2 @ Rf, t2, s2
        thr = min(P[np.argsort(-P[:, 0])[20], 0], 0.5)
        keep = np.where(P[:, 0] > thr)[0]
        dt = self._np_dtype
        self.inlier_A = train_x[keep, :].astype(dt)
        self.inlier_B = train_y[keep, :].astype(dt)
        self.inlier_P = P[keep, :].astype(dt)
        self.init_R = R.astype(dt)
        self.init_t = np.asarray(t).astype(dt)
        if self.init_transform:
            self.inlier_A = self.inlier_A @ self.init_R.T + self.init_t
            self.coordsA = self.coordsA @ self.init_R.T + self.init_t

    def _voxel_data_device(self, coords: np.ndarray, gene_exp: np.ndarray, voxel_size=None, voxel_num: int = 10000):
        """``voxel_data`` (utils.py:1283-1336) on the device: returns (voxel coordinates [n_used, D] numpy, voxel mean
        expression [n_used, G] float64 DEVICE tensor). The grid axes come from the same ``np.arange`` calls as the
        reference (host, tiny); membership uses the reference's test in the coordinates' dtype (csrc/voxel.cu)."""
        dev, lib = self._dev, self._lib
        N, D = coords.shape
        if coords.dtype not in (np.float32, np.float64):
            coords = coords.astype(np.float64)
        lo, hi = np.min(coords, axis=0), np.max(coords, axis=0)
        if voxel_size is None:
            voxel_size = np.sqrt(np.prod(hi - lo)) / (np.sqrt(N) / 5)
        steps = (hi - lo) / int(np.sqrt(voxel_num))
        # np.arange on float32 scalars returns float64, so the reference's `coords - voxel_coord` is evaluated in float64
        axes = [np.ascontiguousarray(np.arange(a, b, st_), dtype=np.float64) for a, b, st_ in zip(lo, hi, steps)]
        grid = np.stack(np.meshgrid(*[np.arange(a, b, st_) for a, b, st_ in zip(lo, hi, steps)]), axis=-1).reshape(-1, D)
        radius = float(voxel_size / 2)
        is_f64 = 1
        cd = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(dev)
        _count_h2d(cd)
        axd = [torch.from_numpy(a).to(dev) for a in axes]
        while len(axd) < 3:
            axd.append(axd[0])
        lo3 = np.zeros(3); lo3[:D] = lo.astype(np.float64)
        st3 = np.ones(3); st3[:D] = np.maximum(steps.astype(np.float64), 1e-300)
        n = [len(a) for a in axes] + [1] * (3 - D)
        counts = torch.zeros((grid.shape[0],), dtype=torch.int32, device=dev)
        stp = _capi.current_stream_ptr()
        geom = (ptr(cd), is_f64, N, D, ptr(axd[0]), n[0], ptr(axd[1]), n[1], ptr(axd[2]), n[2], radius, ptr(lo3), ptr(st3))
        check(lib.spb_voxel_count(*geom, ptr(counts), stp), "spb_voxel_count")
        used = counts > 0
        new_id = (torch.cumsum(used.to(torch.int32), 0) - 1).to(torch.int32)
        n_used = int(used.sum().item())
        if torch.is_tensor(gene_exp):
            ex = gene_exp.to(device=dev, dtype=torch.float32).contiguous()
        else:
            ex = torch.from_numpy(np.ascontiguousarray(gene_exp, dtype=np.float32)).to(dev)
            _count_h2d(ex)
        G = ex.shape[1]
        means = torch.zeros((n_used, G), dtype=torch.float64, device=dev)
        check(lib.spb_voxel_accumulate(*geom, ptr(counts), ptr(new_id), ptr(ex), G, G, ptr(means), G, stp),
              "spb_voxel_accumulate")
        return grid[used.cpu().numpy(), :], means

    def _inlier_from_NN_device(self, train_x: np.ndarray, train_y: np.ndarray, distance: np.ndarray):
        """``inlier_from_NN`` (utils.py:1220-1280) on the device: returns (P [N,1], R [D,D], t [D], sigma2, gamma)."""
        dev, lib = self._dev, self._lib
        N, D = train_x.shape
        x64 = np.zeros((N, 3)); x64[:, :D] = train_x
        y64 = np.zeros((N, 3)); y64[:, :D] = train_y
        dist = np.maximum(0, np.asarray(distance, dtype=np.float64).reshape(-1))
        dist = dist / (np.max(dist) / (np.log(10) * 2))
        area = float(np.maximum(np.prod(train_x.max(0) - train_x.min(0)), np.prod(train_y.max(0) - train_y.min(0))))
        sigma2_init = float(np.sum((train_x.astype(np.float64) - train_y.astype(np.float64)) ** 2) / (D * N))
        w0 = np.exp(-dist)
        xd, yd, dd = (torch.from_numpy(a).to(dev) for a in (x64, y64, dist))
        Pd = torch.from_numpy(w0).to(dev)
        resid = torch.empty((N,), dtype=torch.float64, device=dev)
        state = torch.zeros((128,), dtype=torch.float64, device=dev)
        out = torch.zeros((16,), dtype=torch.float64, device=dev)
        check(
            lib.spb_inlier_from_nn(ptr(xd), ptr(yd), ptr(dd), N, D, area, float(dist.min()), sigma2_init, float(w0.sum()),
                                   ptr(Pd), ptr(resid), ptr(state), ptr(out), _capi.current_stream_ptr()),
            "spb_inlier_from_nn",
        )
        o = out.cpu().numpy()
        R = o[:9].reshape(3, 3)[:D, :D].copy()
        return Pd.cpu().numpy()[:, None], R, o[9 : 9 + D].copy(), float(o[12]), float(o[13])

    # ------------------------------------------------------------------------------------------------------------------
    # variational initialisation (morpho_class.py:683-820; utils.py:1339-1354)
    # ------------------------------------------------------------------------------------------------------------------
    def _init_guess_sigma2(self, subsample: int = 20000) -> float:
        NA, NB, D = self.NA, self.NB, self.D
        sa = np.random.choice(NA, subsample, replace=False) if NA > subsample else np.arange(NA)
        sb = np.random.choice(NB, subsample, replace=False) if NB > subsample else np.arange(NB)
        xa = torch.from_numpy(self.coordsA[sa].astype(np.float64)).to(self._dev)
        xb = torch.from_numpy(self.coordsB[sb].astype(np.float64)).to(self._dev)
        total = torch.zeros((), dtype=torch.float64, device=self._dev)
        for c in range(0, xa.shape[0], 4096):
            d2 = torch.cdist(xa[c : c + 4096], xb) ** 2
            total += (d2 * d2).sum()  # the reference squares the already squared distance (utils.py:1352)
        return float(total.item()) / (D * sa.shape[0] * sa.shape[0])

    def _init_probability_parameters(self, subsample: int = 20000):
        for i, (eA, eB, d_s, p_t, p_p) in enumerate(
            zip(self.exp_layers_A, self.exp_layers_B, self.dissimilarity, self.probability_type, self.probability_parameters)
        ):
            if p_p is not None or p_t.lower() not in ("gauss", "gaussian"):
                continue
            sa = np.random.choice(self.NA, subsample, replace=False) if self.NA > subsample else np.arange(self.NA)
            sb = np.random.choice(self.NB, subsample, replace=False) if self.NB > subsample else np.arange(self.NB)
            if d_s == "label":
                ET, nA = self._raw_cost_T(eA[sa], eB[sb], d_s)
            else:  # gather the sub-samples on the device from the resident copies
                dA, dB = self._to_device_pinned(eA), self._to_device_pinned(eB)
                to_dev = lambda ix: torch.from_numpy(np.ascontiguousarray(ix, dtype=np.int64)).to(self._dev)
                ET, nA = self._raw_cost_T(dA if sa.shape[0] == self.NA else dA.index_select(0, to_dev(sa)),
                                          dB if sb.shape[0] == self.NB else dB.index_select(0, to_dev(sb)), d_s)
            mn = ET[:, :nA].min(dim=0).values  # min over fixed cells for every moving cell (utils: nx.min(exp_dist, 1))
            srt = torch.sort(mn).values
            val = float(srt[int(sa.shape[0] * 0.05)].item()) / 5
            self.probability_parameters[i] = np.maximum(np.asarray(val, dtype=self._np_dtype), np.asarray(0.01, dtype=self._np_dtype))
            del ET

    def _initialize_variational_variables(self):
        dt = self._np_dtype
        self.sigma2 = np.asarray(self.sigma2_init_scale * self._init_guess_sigma2(), dtype=dt)
        self._sigma2_init = float(self.sigma2)
        self._init_probability_parameters()
        self.sigma2_variance = 1.0
        self.sigma2_variance_end = float(self.partial_robust_level)
        self.sigma2_variance_decress = float(np.power(np.asarray(self.sigma2_variance_end / self.sigma2_variance, dtype=dt), 1 / 100))
        if isinstance(self.kappa, float):
            self.kappa = np.ones((self.NA,), dtype=dt) * self.kappa
        elif isinstance(self.kappa, np.ndarray):
            self.kappa = self.kappa.astype(dt)
        else:
            raise ValueError("kappa should be a float or a numpy array.")
        self.gamma = np.asarray(0.5, dtype=dt)
        self.samples_s = float(
            np.maximum(
                np.prod(self.coordsA.max(axis=0) - self.coordsA.min(axis=0)),
                np.prod(self.coordsB.max(axis=0) - self.coordsB.min(axis=0)),
            )
        )
        self.outlier_s = self.samples_s * self.NA
        self.nonrigid_flag = False
        if self.SVI_mode:
            if self.batch_size is None:
                self.batch_size = min(max(int(self.NB / 10), 1000), self.NB)
            else:
                self.batch_size = min(self.batch_size, self.NB)
            self.batch_perm = np.random.permutation(self.NB)

    # ------------------------------------------------------------------------------------------------------------------
    # device state
    # ------------------------------------------------------------------------------------------------------------------
    def _set_row_order(self):
        """Processing order of the moving cells (after the coarse initialisation moved them)."""
        if self.spatial_sort and self.NA > _capi.ROW_TILE:
            self._perm = morton_order(self.coordsA)
            self._perm_dev = torch.from_numpy(self._perm).to(self._dev)
        else:
            self._perm, self._perm_dev = None, None

    def _sorted(self, host_rows: np.ndarray) -> np.ndarray:
        return host_rows if self._perm is None else host_rows[self._perm]

    def _unsorted(self, sorted_rows: np.ndarray) -> np.ndarray:
        if self._perm is None:
            return sorted_rows
        out = np.empty_like(sorted_rows)
        out[self._perm] = sorted_rows
        return out

    def _build_gene_cost(self):
        """GT[j][i] = prod_layers prob(metric(A_i, B_j)) (morpho_class.py:265-268 + utils.py:1080-1081)."""
        dev, lib = self._dev, self._lib
        self._set_row_order()
        c0, c1 = self._col_range()
        nb_loc = c1 - c0
        self._GT = torch.empty((nb_loc, self.ldx), dtype=torch.float32, device=dev)
        gc = GeneCostBuilder(lib, dev)
        first = True
        for eA, eB, d_s, p_t, p_p in zip(
            self.exp_layers_A, self.exp_layers_B, self.dissimilarity, self.probability_type, self.probability_parameters
        ):
            if d_s == "label":
                la = torch.from_numpy(np.ascontiguousarray(eA if self._perm is None else eA[self._perm], dtype=np.int32)).to(dev)
                lb = torch.from_numpy(np.ascontiguousarray(eB[c0:c1], dtype=np.int32)).to(dev)
                LT = torch.from_numpy(np.ascontiguousarray(self.label_transfer, dtype=np.float32)).to(dev)
                check(
                    lib.spb_label_cost(ptr(la), ptr(lb), ptr(LT), LT.shape[1], self.NA, nb_loc, 0 if first else 1,
                                       ptr(self._GT), self.ldx, _capi.current_stream_ptr()),
                    "spb_label_cost",
                )
            else:
                A = self._to_device_pinned(eA)
                if self._perm is not None:
                    A = A.index_select(0, self._perm_dev)  # moving cells in Morton order
                B = self._to_device_pinned(eB) if self.column_shard is None else staged_to_device(eB[c0:c1], dev)
                opA, rtA, opB, rtB, G_eff = gc.prepare_pair(A, B, d_s)
                gc.cost(opA, rtA, opB, rtB, self.NA, nb_loc, G_eff, d_s, p_t, p_p, not first, self._GT, self.ldx)
                del A, B, opA, opB
            first = False
        self.__dict__.pop("_dev_rep", None)  # the resident copies of the representations are no longer needed

    def _col_range(self):
        """Fixed cells (columns of P) held by this process: all of them, or this rank's block of a column-sharded pair."""
        if self.column_shard is None:
            return 0, self.NB
        r, w = int(self.column_shard[0]), int(self.column_shard[1])
        return (self.NB * r) // w, (self.NB * (r + 1)) // w

    def _to_device_pinned(self, host_array: np.ndarray) -> torch.Tensor:
        """Device copy of one dense representation, uploaded ONCE per preparation (the coarse initialisation, the beta^2
        initialisation and the cost matrix all read it): straight from pinned memory when ``pin_inputs`` staged it,
        otherwise through the reusable pinned staging buffers."""
        key = id(host_array)
        dcache = self.__dict__.setdefault("_dev_rep", {})
        if key in dcache:
            return dcache[key]
        cache = self.__dict__.setdefault("_pinned", {})
        self._h2d_bytes = getattr(self, "_h2d_bytes", 0) + host_array.size * 4
        if key in cache:
            _count_h2d(cache[key])
            t = cache[key].to(self._dev, non_blocking=True)
        else:
            t = staged_to_device(host_array, self._dev)
        dcache[key] = t
        return t

    def _device_rows(self, layer: str, field: str, side: str, idx: np.ndarray):
        """Rows ``idx`` of a dense representation as a device tensor, gathered on the device when the representation is
        one of the alignment's own layers (the usual case: init_layer == rep_layer); None when it is not."""
        for r, f, eA, eB in zip(self.rep_layer, self.rep_field, self.exp_layers_A, self.exp_layers_B):
            if r == layer and f == field and f != "obs":
                full = self._to_device_pinned(eA if side == "A" else eB)
                if idx.shape[0] == full.shape[0] and np.array_equal(idx, np.arange(full.shape[0])):
                    return full
                return full.index_select(0, torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(self._dev))
        return None

    def pin_inputs(self):
        """Stage the dense representations in pinned host memory ahead of time (part of preprocessing)."""
        for e in list(self.exp_layers_A) + list(self.exp_layers_B):
            if e.dtype.kind == "f":
                key = id(e)
                cache = self.__dict__.setdefault("_pinned", {})
                if key not in cache:
                    t = torch.from_numpy(np.ascontiguousarray(e, dtype=np.float32))
                    cache[key] = t if t.is_pinned() else t.pin_memory()

    @staticmethod
    def _choose_segments(nrb: int, nbb: int) -> int:
        """Column segments so that CTAs ~ a multiple of the resident CTA slots (148 SMs x 2048 / ROW_TILE) and a segment is at most ``SPB_MAX_COLS_PER_CTA`` columns
        (short CTAs keep the tail of the last wave small once culling has shortened the column lists)."""
        cap = int(os.environ.get("SPB_MAX_COLS_PER_CTA", "4096"))
        max_seg = max(1, nbb // _capi.COL_STAGE)
        slots = 148 * (2048 // _capi.ROW_TILE)  # resident CTAs of the sweeps on one B200
        seg = 1
        for waves in range(1, 256):
            seg = max(1, min(max_seg, (slots * waves) // max(nrb, 1)))
            if (nbb + seg - 1) // seg <= cap or seg == max_seg:
                break
        return seg

    def _allocate_state(self):
        dev, D, NA, K, ldx = self._dev, self.D, self.NA, self.K, self.ldx
        f32, f64 = torch.float32, torch.float64
        c0, c1 = self._col_range()
        NB = c1 - c0  # columns held by this process (all of them unless the pair is column-sharded)
        if self.column_shard is not None and (self.SVI_mode or self.sparse_calculation_mode or self.guidance or self.materialize_P
                                              or self.compute_mapping or self.return_mapping):
            raise NotImplementedError("a column-sharded pair supports the full EM only (SVI_mode=False, materialize_P=False, no "
                                      "sparse mode / guidance / mapping outputs)")
        nbb = self.batch_size if self.SVI_mode else NB
        nbb_alloc = NB if (self.return_mapping and self.SVI_mode) else nbb
        self._NBb = nbb
        nrb = ldx // _capi.ROW_TILE
        self._nbb_pad = _round_up(nbb_alloc, 8) + 8
        s = {}
        s["xa"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["xa"][:D, :NA] = torch.from_numpy(np.ascontiguousarray(self._sorted(self.coordsA).T, dtype=np.float32)).to(dev)
        s["xb4"] = torch.zeros((NB, 4), dtype=f32, device=dev)
        s["xb4"][:, :D] = torch.from_numpy(self.coordsB[c0:c1].astype(np.float32)).to(dev)
        s["Gamma"] = torch.from_numpy(np.ascontiguousarray(self.GammaSparse, dtype=np.float32)).to(dev)
        s["kappa"] = torch.ones((ldx,), dtype=f32, device=dev)
        s["kappa"][:NA] = torch.from_numpy(self._sorted(self.kappa).astype(np.float32)).to(dev)
        if self._perm is not None and not getattr(self, "_UT_is_sorted", False):
            ut = torch.zeros_like(self._UT)
            ut[:, :NA] = self._UT[:, :NA].index_select(1, self._perm_dev)
            self._UT, self._UT_is_sorted = ut, True
        s["alpha"] = torch.ones((ldx,), dtype=f32, device=dev)
        s["SigmaDiag"] = torch.zeros((ldx,), dtype=f32, device=dev)
        s["lm"] = torch.zeros((ldx,), dtype=f32, device=dev)
        s["mm"] = torch.zeros((ldx,), dtype=f32, device=dev)
        s["VnA"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["RnA"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["XAHat"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["XAHat"][:, NA:] = 1e18  # pad rows sit infinitely far from every fixed cell
        for k in ("K_NA", "K_NA_spatial", "K_NA_sigma2"):
            s[k] = torch.zeros((ldx,), dtype=f32, device=dev)
        s["PXB"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["PXB_term"] = torch.zeros((3, ldx), dtype=f32, device=dev)
        s["K_NB"] = torch.zeros((self._nbb_pad,), dtype=f32, device=dev)
        s["colgeom"] = torch.zeros((self._nbb_pad, 8), dtype=f32, device=dev)
        s["colconst"] = torch.zeros((self._nbb_pad, _capi.CONST["SPB_COLCONST_FLOATS"]), dtype=f32, device=dev)
        s["colpart"] = torch.zeros((nrb, 4, self._nbb_pad), dtype=f32, device=dev)
        s["keepmask"] = torch.zeros((nrb, (self._nbb_pad + 31) // 32), dtype=torch.int32, device=dev)
        seg1 = self._choose_segments(nrb, nbb)
        seg2 = self._choose_segments(nrb, nbb)
        seg_alloc = max(seg2, self._choose_segments(nrb, nbb_alloc))
        s["rowpart"] = torch.zeros((seg_alloc, 8, ldx), dtype=f32, device=dev)
        s["bbox"] = torch.zeros((nrb, 8), dtype=f32, device=dev)
        s["collist"] = torch.zeros((nrb, self._nbb_pad), dtype=torch.int32, device=dev)
        s["colcount"] = torch.zeros((nrb,), dtype=torch.int32, device=dev)
        s["colsplit"] = torch.zeros((nrb,), dtype=torch.int32, device=dev)
        if self.sparse_calculation_mode:
            s["colmask"] = torch.zeros((self._nbb_pad, _capi.CONST["SPB_COLMASK_WORDS"]), dtype=torch.int32, device=dev)
        s["UtWU"] = torch.zeros((K, K), dtype=f64, device=dev)
        s["UtPXB"] = torch.zeros((K, 3), dtype=f64, device=dev)
        s["SigmaInv"] = torch.zeros((K, K), dtype=f64, device=dev)
        s["Sigma"] = torch.zeros((K, K), dtype=f64, device=dev)
        s["Coff"] = torch.zeros((K, 3), dtype=f64, device=dev)
        s["moments"] = torch.zeros((32,), dtype=f64, device=dev)
        s["trace_buf"] = torch.zeros((max(self.max_iter, 1), _capi.TRACE_STRIDE), dtype=f64, device=dev)
        s["optimal"] = torch.zeros((12,), dtype=f64, device=dev)
        if self.SVI_mode:
            # the whole batch schedule is a deterministic function of the initial permutation (morpho_class.py:894-896)
            sched = np.empty((max(self.max_iter, 1), nbb), dtype=np.int32)
            perm = self.batch_perm.copy()
            for it in range(self.max_iter):
                sched[it] = perm[:nbb]
                perm = np.roll(perm, nbb)
            self.batch_idx = sched[self.max_iter - 1].astype(np.int64) if self.max_iter > 0 else None
            s["batch_idx"] = torch.from_numpy(sched).to(dev)
        else:
            s["batch_idx"] = None
        # scalars
        sc = SpbScalars()
        sc.sigma2 = float(self._sigma2_init)
        sc.sigma2_variance = 1.0
        sc.gamma = 0.5
        sc.iter = -1  # device-side iteration counter: graph replays advance it (spb_em_iteration_ex)
        for q in range(9):
            sc.R[q] = 1.0 if q in (0, 4, 8) else 0.0
        host_sc = np.frombuffer(bytes(sc), dtype=np.uint8).copy()
        s["sc"] = torch.from_numpy(host_sc).to(dev)
        self._state = s
        self.__dict__.pop("_graphs", None)  # captured iteration graphs hold the old state's pointers
        # params
        p = SpbEmParams()
        p.NA, p.NB, p.NBb, p.D, p.K, p.ldx = NA, NB, nbb, D, K, ldx
        p.NB_total = self.NB if self.column_shard is not None else 0
        # block partials + tickets of the ordered (reproducible) grid reductions
        n_red = max(592 * 29, ((NA + 255) // 256) * 4, 320 * K * (K + 3) if K <= 32 else 0) + 64
        s["red_scratch"] = torch.zeros((n_red,), dtype=f64, device=dev)
        s["red_counter"] = torch.zeros((8,), dtype=torch.int32, device=dev)
        p.red_scratch, p.red_scratch_doubles, p.red_counter = s["red_scratch"].data_ptr(), n_red, s["red_counter"].data_ptr()
        p.shard_rank = p.shard_world = 0
        p.rowstat = p.peer_rowstat = p.shard_flags = p.peer_flags = None
        if self.column_shard is not None:
            self._setup_column_shard(p, s)
        p.svi, p.nn_init, p.update_R = int(self.SVI_mode), int(self.nn_init), int(self.update_R)
        p.nonrigid_start_iter = int(self.nonrigid_start_iter)
        p.seg1, p.seg2, p.nbb_pad, p.trace = seg1, seg2, self._nbb_pad, 1
        p.cull = int(bool(self.cull_zero_tiles))
        p.sparse_k = int(self.sparse_top_k) if self.sparse_calculation_mode else 0
        p.lambdaVF, p.gamma_a, p.gamma_b = float(self.lambdaVF), float(self.gamma_a), float(self.gamma_b)
        p.samples_s = float(self.samples_s)
        p.pinv_eps = self._pinv_eps()
        p.nn_init_weight = float(self.nn_init_weight)
        p.sigma2_variance_decress = float(self.sigma2_variance_decress)
        p.sigma2_variance_end = float(self.sigma2_variance_end)
        if self.nn_init:
            Pn = self.inlier_P.astype(np.float64)[:, 0]
            a = np.zeros((Pn.shape[0], 3))
            b = np.zeros((Pn.shape[0], 3))
            a[:, :D] = self.inlier_A.astype(np.float64)
            b[:, :D] = self.inlier_B.astype(np.float64)
            p.inl_SP = float(Pn.sum())
            Sa, Sb = Pn @ a, Pn @ b
            Mab = (a * Pn[:, None]).T @ b
            for d in range(3):
                p.inl_Sa[d], p.inl_Sb[d] = float(Sa[d]), float(Sb[d])
            for q in range(9):
                p.inl_Mab[q] = float(Mab.reshape(-1)[q])
        else:
            p.inl_SP = 1.0
        if self.guidance:
            NI = self.X_AI.shape[0]
            pad = lambda a: np.pad(np.asarray(a, dtype=np.float64), ((0, 0), (0, 3 - a.shape[1])))
            s["g_XA"] = torch.from_numpy(pad(self.X_AI)).to(dev)
            s["g_XB"] = torch.from_numpy(pad(self.X_BI)).to(dev)
            s["g_VA"] = torch.zeros((NI, 3), dtype=f64, device=dev)
            s["g_RA"] = torch.zeros((NI, 3), dtype=f64, device=dev)
            UI = self.U_I if self.U_I is not None else np.zeros((NI, K))
            s["g_UI"] = torch.from_numpy(np.ascontiguousarray(UI, dtype=np.float64)).to(dev)
            s["g_G1"] = torch.from_numpy(np.ascontiguousarray(UI.T @ UI, dtype=np.float64)).to(dev)
            p.g_on, p.g_NI = 1, NI
            p.g_nonrigid = int(self.guidance_effect in ("nonrigid", "both"))
            p.g_rigid = int(self.guidance_effect in ("rigid", "both"))
            p.g_weight = float(self.guidance_weight)
            p.g_meanXB, p.g_meanXA = float(self.X_BI.astype(np.float64).mean()), float(self.X_AI.astype(np.float64).mean())
            for name in ("g_XA", "g_XB", "g_VA", "g_RA", "g_UI", "g_G1"):
                setattr(p, name, s[name].data_ptr())
        p.GT, p.UT = ptr(self._GT).value, ptr(self._UT).value
        for name in ("xa", "xb4", "Gamma", "kappa", "batch_idx", "alpha", "SigmaDiag", "lm", "mm", "VnA", "RnA", "XAHat",
                     "K_NA", "K_NA_spatial", "K_NA_sigma2", "PXB", "PXB_term", "K_NB", "colgeom", "colconst", "colpart", "keepmask",
                     "rowpart", "bbox", "collist", "colcount", "colsplit", "UtWU", "UtPXB", "SigmaInv", "Sigma", "Coff", "moments", "sc",
                     "trace_buf"):
            t = s[name]
            setattr(p, name, None if t is None else t.data_ptr())
        # eigenbasis of the previous non-rigid solve (warm start of the in-library Jacobi); [0] = K once valid
        s["jacobi_ws"] = torch.zeros((1 + K * K,), dtype=f64, device=dev) if K <= _capi.MAX_K_FUSED else None
        p.jacobi_ws = None if s["jacobi_ws"] is None else s["jacobi_ws"].data_ptr()
        p.colmask = s["colmask"].data_ptr() if "colmask" in s else None
        # K^T P K contraction: tcgen05 (3xTF32) above 32 inducing points, exact fp64 SIMT kernel for small K
        backend = os.environ.get("SPB_GRAM", "auto")
        if backend == "tensor" or (backend == "auto" and K > 32):
            if "UT_hi" not in self.__dict__.setdefault("_gram", {}) or self._gram["UT_hi"].shape != self._UT.shape:
                hi, lo = torch.empty_like(self._UT), torch.empty_like(self._UT)
                mean = torch.empty((K,), dtype=f32, device=dev)
                check(self._lib.spb_gram_center(ptr(self._UT), ldx, NA, K, ptr(mean), ptr(hi), ptr(lo), _capi.current_stream_ptr()),
                      "spb_gram_center")
                need = C.c_int64(0)
                check(self._lib.spb_gram_tc_scratch_floats(K, 3, NA, C.byref(need)), "spb_gram_tc_scratch_floats")
                self._gram = dict(
                    UT_hi=hi, UT_lo=lo, mean=mean,
                    GB_hi=torch.zeros((K + 4, ldx), dtype=f32, device=dev), GB_lo=torch.zeros((K + 4, ldx), dtype=f32, device=dev),
                    scratch=torch.empty((need.value,), dtype=f32, device=dev), sums=torch.zeros((4,), dtype=f64, device=dev),
                )
            g = self._gram
            p.UT_hi, p.UT_lo, p.UT_mean = g["UT_hi"].data_ptr(), g["UT_lo"].data_ptr(), g["mean"].data_ptr()
            p.GB_hi, p.GB_lo, p.gram_sums = g["GB_hi"].data_ptr(), g["GB_lo"].data_ptr(), g["sums"].data_ptr()
            p.gram_scratch, p.gram_scratch_floats = g["scratch"].data_ptr(), g["scratch"].numel()
        else:
            p.UT_hi = p.UT_lo = p.UT_mean = p.GB_hi = p.GB_lo = p.gram_scratch = p.gram_sums = None
            p.gram_scratch_floats = 0
        self._params = p

    def _setup_column_shard(self, p, s):
        """Buffers of the column-sharded pair: fp64 row statistics of this rank's columns (double-buffered) + epoch flags.
        ``mode`` "p2p" (default when available): the buffer is symmetric memory, every rank maps every peer's copy and the
        row-finalize kernel sums them straight over NVLink; "nccl": plain buffer + ``all_reduce`` (the baseline)."""
        import torch.distributed as dist

        rank, world = int(self.column_shard[0]), int(self.column_shard[1])
        mode = self.column_shard[2] if len(self.column_shard) > 2 else "auto"
        dev, ldx = self._dev, self.ldx
        n_stat = 2 * 8 * ldx
        p.shard_rank, p.shard_world = rank, world
        self._shard_epoch = 0
        self._shard_mode = "nccl"
        buf = None
        if mode in ("auto", "p2p") and world > 1:
            try:
                import torch.distributed._symmetric_memory as symm

                buf = symm.empty((n_stat + 64,), dtype=torch.float64, device=dev)
                buf.zero_()
                try:
                    hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
                except Exception:
                    hdl = symm.rendezvous(buf, dist.group.WORLD)
                ptrs = [int(q) for q in hdl.buffer_ptrs]
                s["shard_hdl"] = hdl
                s["peer_rowstat"] = torch.tensor(ptrs, dtype=torch.int64, device=dev)
                s["peer_flags"] = torch.tensor([q + n_stat * 8 for q in ptrs], dtype=torch.int64, device=dev)
                p.peer_rowstat, p.peer_flags = s["peer_rowstat"].data_ptr(), s["peer_flags"].data_ptr()
                self._shard_mode = "p2p"
                hdl.barrier()
            except Exception as e:  # no symmetric memory on this system: fall back to the collective
                if mode == "p2p":
                    raise
                buf = None
                self._shard_fallback_reason = repr(e)
        if buf is None:
            buf = torch.zeros((n_stat + 64,), dtype=torch.float64, device=dev)
        s["rowstat"] = buf
        p.rowstat = buf.data_ptr()
        p.shard_flags = buf.data_ptr() + n_stat * 8

    def _shard_row_statistics(self, st):
        """Row statistics of a column-sharded pair: local fold, sum over the ranks, finish (replaces spb_row_finalize)."""
        import torch.distributed as dist

        lib, p, s = self._lib, self._params, self._state
        parity = self._shard_epoch & 1
        self._shard_epoch += 1
        check(lib.spb_row_fold(C.byref(p), parity, st), "spb_row_fold")
        if self._shard_mode == "p2p":
            check(lib.spb_row_stats_p2p(C.byref(p), parity, self._shard_epoch, st), "spb_row_stats_p2p")
        else:
            view = s["rowstat"][parity * 8 * self.ldx : (parity + 1) * 8 * self.ldx]
            hook = getattr(self, "_shard_reduce_hook", None)
            if hook is not None:  # tests: several shards stepped in lock-step inside one process
                hook(self, view)
            elif dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(view)
            check(lib.spb_row_stats_finalize(C.byref(p), parity, st), "spb_row_stats_finalize")

    def _pinv_eps(self) -> float:
        """Machine epsilon behind scipy.linalg.pinv's default cutoff in the reference (utils.py:1435): float32 for the
        Euclidean kernel; the geodesic kernel matrix is float64 there (con_K_graph, utils.py:1208-1217), which promotes
        SigmaInv and makes the cutoff K * eps(float64)."""
        return 2.220446049250313e-16 if self.kernel_type == "geodist" else 1.1920928955078125e-07

    def _read_scalars(self) -> SpbScalars:
        raw = self._state["sc"].cpu().numpy().tobytes()
        return SpbScalars.from_buffer_copy(raw)

    # ------------------------------------------------------------------------------------------------------------------
    # the EM loop
    # ------------------------------------------------------------------------------------------------------------------
    def _nonrigid_solve_large_K(self, st):
        """K > SPB_MAX_K_FUSED: eigen pseudo-inverse through cuSOLVER (torch.linalg.eigh, fp64), same cutoff rule as
        scipy.linalg.pinv on the reference's fp32 matrix (utils.py:1435). Everything stays on the device: the kept
        eigen-directions are sorted first and handed to the row kernel as a factor of Sigma, with their count in device memory."""
        lib, p, s = self._lib, self._params, self._state
        check(lib.spb_nonrigid_blend(C.byref(p), st), "spb_nonrigid_blend")
        A = s["SigmaInv"]
        A = 0.5 * (A + A.T)
        ev, V = torch.linalg.eigh(A)
        ev, V = ev.flip(0), V.flip(1)  # descending: directions above the cutoff come first
        cutoff = ev.abs().max() * self.K * self._pinv_eps()
        keep = ev.abs() > cutoff
        inv = torch.where(keep, 1.0 / ev, torch.zeros_like(ev))
        VS = V * inv
        s["Sigma"].copy_(VS @ V.T)
        rhs = s["UtPXB"]
        g_nonrigid = self.guidance and self.guidance_effect in ("nonrigid", "both")
        if g_nonrigid:  # morpho_class.py:1286-1288, 1294-1295 (the SigmaInv part is added by spb_nonrigid_blend)
            sc = s["sc"][:80].view(torch.float64)  # sigma2 = [0], Sp = [3] (spb_scalars layout)
            cg = sc[0] * float(self.guidance_weight) * sc[3] / self.X_AI.shape[0]
            rhs = rhs + cg * (s["g_UI"].T @ (s["g_XB"] - s["g_RA"]))
        s["Coff"].copy_(VS @ (V.T @ rhs))
        if g_nonrigid:
            s["g_VA"].copy_(s["g_UI"] @ s["Coff"])
        # factor of Sigma for the row kernel: G = V sqrt(|inv|) sign-safe (kept eigenvalues of the PSD matrix are positive)
        G = s.setdefault("sigma_factor", torch.zeros((self.K, self.K), dtype=torch.float64, device=self._dev))
        G.copy_(V * torch.sqrt(inv.clamp_min(0.0)))
        rank = s.setdefault("sigma_rank", torch.zeros((1,), dtype=torch.int32, device=self._dev))
        rank.copy_((keep & (ev > 0)).sum().to(torch.int32).reshape(1))

    def _iteration(self, it: int, st, capture_P: bool = False, sweep_events: Optional[list] = None):
        """One EM iteration (morpho_class.py:280-294). The fused C entry point is used unless the iteration has to be
        split: K > SPB_MAX_K_FUSED (eigen solve through cuSOLVER) or ``capture_P`` (dense P of THIS E-step, which must
        be written before the M-step moves the cells)."""
        lib, p = self._lib, self._params
        nonrigid = it > self.nonrigid_start_iter
        large_K = nonrigid and self.K > _capi.MAX_K_FUSED
        if not (large_K or capture_P or sweep_events is not None or self.column_shard is not None):
            check(lib.spb_em_iteration(C.byref(p), it, st), "spb_em_iteration")
            return
        self._estep_only(it, st, sweep_events)
        if capture_P:
            self._capture_P(it, st)
        check(lib.spb_update_gamma_alpha(C.byref(p), st), "spb_update_gamma_alpha")
        if nonrigid:
            check(lib.spb_nonrigid_accumulate(C.byref(p), st), "spb_nonrigid_accumulate")
            if large_K:
                self._nonrigid_solve_large_K(st)
                check(lib.spb_field_apply_lowrank(C.byref(p), ptr(self._state["sigma_factor"]), self.K,
                                                  ptr(self._state["sigma_rank"]), st), "spb_field_apply_lowrank")
            else:
                check(lib.spb_nonrigid_solve(C.byref(p), st), "spb_nonrigid_solve")
                check(lib.spb_field_apply(C.byref(p), st), "spb_field_apply")
        check(lib.spb_rigid_moments(C.byref(p), st), "spb_rigid_moments")
        check(lib.spb_rigid_solve(C.byref(p), it, st), "spb_rigid_solve")
        check(lib.spb_row_update(C.byref(p), st), "spb_row_update")

    def _capture_P(self, it: int, st):
        """Posterior of the E-step that has just run: dense [N_A, NBb], or in sparse_calculation_mode the COO entries
        (top-k rows and values per column) without ever forming the dense matrix."""
        lib, p = self._lib, self._params
        if self.compute_mapping:  # row / column maxima of the same posterior, straight from the cost matrix
            self._rowbest = torch.zeros((self.NA,), dtype=torch.int64, device=self._dev)
            self._colbest = torch.zeros((self._NBb,), dtype=torch.int64, device=self._dev)
            check(lib.spb_posterior_argmax(C.byref(p), it, ptr(self._rowbest), ptr(self._colbest), st), "spb_posterior_argmax")
        if not self.materialize_P:
            self._P_dev = "skipped"
            return
        if self.sparse_calculation_mode:
            k = int(self.sparse_top_k)
            self._P_rows = torch.zeros((self._NBb, k), dtype=torch.int32, device=self._dev)
            self._P_vals = torch.zeros((self._NBb, k), dtype=torch.float32, device=self._dev)
            check(lib.spb_sparse_P_emit(C.byref(p), it, ptr(self._P_rows), ptr(self._P_vals), st), "spb_sparse_P_emit")
            self._P_dev = "sparse"
        else:
            self._P_dev = torch.empty((self.NA, self._NBb), dtype=torch.float32, device=self._dev)
            check(lib.spb_materialize_P(C.byref(p), it, ptr(self._P_dev), self._NBb, st), "spb_materialize_P")

    def _sparse_P_to_coo(self, dt):
        """scipy COO in the reference's layout (utils.py:1385-1392,1506-1510): per column the k entries in descending
        order, columns concatenated."""
        import scipy.sparse as sp

        k = min(int(self.sparse_top_k), self.NA)
        vals, order = torch.sort(self._P_vals[:, :k], dim=1, descending=True, stable=True)
        rows = torch.gather(self._P_rows[:, :k].long(), 1, order).cpu().numpy()
        if self._perm is not None:
            rows = self._perm[rows]
        col = np.repeat(np.arange(self._NBb), k)
        return sp.coo_matrix((vals.cpu().numpy().astype(dt).reshape(-1), (rows.reshape(-1), col)),
                             shape=(self.NA, self._NBb))

    def _estep_only(self, it: int, st, sweep_events: Optional[list] = None):
        """One E-step + the statistics the closing similarity needs (used for return_mapping under SVI)."""
        lib, p = self._lib, self._params
        check(lib.spb_iter_begin(C.byref(p), it, st), "spb_iter_begin")
        check(lib.spb_gather_cols(C.byref(p), it, st), "spb_gather_cols")
        check(lib.spb_estep_col_lists(C.byref(p), st), "spb_estep_col_lists")
        if sweep_events is not None:
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
        check(lib.spb_estep_sweep1(C.byref(p), it, st), "spb_estep_sweep1")
        if sweep_events is not None:
            e1.record()
        check(lib.spb_col_finalize(C.byref(p), st), "spb_col_finalize")
        if self.sparse_calculation_mode:
            check(lib.spb_estep_col_select(C.byref(p), it, st), "spb_estep_col_select")
        if sweep_events is not None:
            e2.record()
        check(lib.spb_estep_sweep2(C.byref(p), it, st), "spb_estep_sweep2")
        if sweep_events is not None:
            e3.record()
            sweep_events.append((e0, e1, e2, e3))
        if self.column_shard is not None:
            self._shard_row_statistics(st)
        else:
            check(lib.spb_row_finalize(C.byref(p), st), "spb_row_finalize")

    def prepare_host(self):
        """Coarse rigid initialisation + variational initialisation (host numpy with small device helpers); consumes
        the global ``np.random`` stream in the reference's order (morpho_class.py:258-261)."""
        import time as _time

        self._timing = {}
        with torch.cuda.device(self._dev):
            t0 = _time.perf_counter()
            if self.nn_init:
                with _nvtx("coarse_rigid_alignment"):
                    self._coarse_rigid_alignment()
            # (stream-level waits: a second pair may be running its EM on another stream of this device)
            torch.cuda.current_stream().synchronize()
            self._timing["coarse_rigid_alignment_s"] = _time.perf_counter() - t0
            t0 = _time.perf_counter()
            self._initialize_variational_variables()
            torch.cuda.current_stream().synchronize()
            self._timing["variational_init_s"] = _time.perf_counter() - t0
        self._host_ready = True

    def prepare_device(self):
        """Host -> device copies of the inputs, expression-cost matrix, EM state (morpho_class.py:265-268)."""
        if not getattr(self, "_host_ready", False):
            self.prepare_host()
        with torch.cuda.device(self._dev):
            with _nvtx("expression_cost_matrix"):
                self._build_gene_cost()
            self._allocate_state()
            check(self._lib.spb_row_update(C.byref(self._params), _capi.current_stream_ptr()), "spb_row_update")
        self._prepared = True

    def prepare(self):
        """Everything ``run`` does before the loop: coarse init, variational init, cost matrix, device state."""
        self.prepare_host()
        self.prepare_device()

    def reset_state(self):
        """Rewind the EM state to iteration 0 (benchmark repetitions); the cost matrix stays resident."""
        with torch.cuda.device(self._dev):
            self._allocate_state()
            check(self._lib.spb_row_update(C.byref(self._params), _capi.current_stream_ptr()), "spb_row_update")

    def run_em(self, n_iter: Optional[int] = None, start: int = 0, sweep_events: Optional[list] = None):
        """Enqueue EM iterations [start, start + n_iter) on the current stream (no host synchronisation).

        Runs of iterations of the same phase (rigid-only up to ``nonrigid_start_iter``, with the non-rigid solve after) are
        replayed from ONE captured CUDA graph of the iteration's launch sequence (``use_cuda_graph``, default on; the
        iteration index — SVI batch, step size, trace row — is a device counter). Iterations that need host involvement
        (history recording, per-sweep events, the posterior capture of the last iteration, K > SPB_MAX_K_FUSED in the
        non-rigid phase) take the plain path.

        ``sweep_events``: optional list that receives (start, mid, end) CUDA events recorded around the two E-step sweep
        kernels of every iteration on the launching stream (bench.py's live roofline measurement)."""
        n_iter = self.max_iter - start if n_iter is None else n_iter
        with torch.cuda.device(self._dev):
            st = _capi.current_stream_ptr()
            hist = self._state.get("hist")
            it, end = start, start + n_iter
            while it < end:
                last = it == self.max_iter - 1
                want_P = (self.materialize_P or self.compute_mapping) and last and not (self.return_mapping and self.SVI_mode)
                nonrigid = it > self.nonrigid_start_iter
                plain = (hist is not None or sweep_events is not None or want_P or _nvtx.enabled or self.column_shard is not None
                         or not getattr(self, "use_cuda_graph", True) or (nonrigid and self.K > _capi.MAX_K_FUSED))
                if not plain:
                    # iterations [it, stop) share the phase and need nothing from the host
                    stop = min(end, self.nonrigid_start_iter + 1) if not nonrigid else end
                    if (self.materialize_P or self.compute_mapping) and stop == self.max_iter and not (self.return_mapping and self.SVI_mode):
                        stop -= 1  # the last iteration captures the posterior: plain path
                    if stop - it >= 3:
                        self._iteration(it, st)  # explicit index: also the warm-up launch of every kernel of the phase
                        if not nonrigid and end > self.nonrigid_start_iter + 4 and self.K <= _capi.MAX_K_FUSED:
                            # capture the graph of the later non-rigid phase now as well: a capture synchronises the
                            # device, and doing it here keeps the host free to enqueue the whole run without stopping
                            check(self._lib.spb_nonrigid_warm(), "spb_nonrigid_warm")
                            self._iteration_graph(True)
                            if self._graph_unroll() > 1 and end - self.nonrigid_start_iter - 2 >= 2 * self._graph_unroll():
                                self._iteration_graph(True, self._graph_unroll())
                        n_rep = stop - it - 1
                        unroll = self._graph_unroll()
                        replayed = 0
                        if unroll > 1 and n_rep >= 2 * unroll:
                            graph_u, n_kernels_u = self._iteration_graph(nonrigid, unroll)
                            for _ in range(n_rep // unroll):
                                graph_u.replay()
                            replayed += n_kernels_u * (n_rep // unroll)
                            n_rep -= (n_rep // unroll) * unroll
                        graph, n_kernels = self._iteration_graph(nonrigid)
                        for _ in range(n_rep):
                            graph.replay()
                        # kernels launched through graph replays are not seen by the library's launch counter
                        self.graph_replayed_launches = getattr(self, "graph_replayed_launches", 0) + replayed + n_kernels * n_rep
                        it = stop
                        continue
                if hist is not None:
                    hist[it].copy_(self._state["XAHat"])
                    self._state["hist_sigma2"][it].copy_(self._state["sc"][:8].view(torch.float64)[0])
                if _nvtx.enabled:
                    torch.cuda.nvtx.range_push(f"em_iteration_{it}")
                self._iteration(it, st, capture_P=want_P, sweep_events=sweep_events)
                if _nvtx.enabled:
                    torch.cuda.nvtx.range_pop()
                it += 1

    def _graph_unroll(self) -> int:
        """Iterations per captured graph: 8 when an iteration touches fewer than 2e9 cell pairs (about 2 ms of device time:
        the default SVI batch of the 100k pair, or any small pair), where one graph launch per iteration can make a slow
        host the bottleneck; 1 for the heavy full-EM iterations."""
        u = getattr(self, "graph_unroll", 0)
        if u > 0:
            return u
        cols = self.batch_size if self.SVI_mode else self.NB
        return 8 if float(self.NA) * float(cols) < 2e9 else 1

    def _iteration_graph(self, nonrigid: bool, unroll: int = 1):
        """CUDA graph of ``unroll`` consecutive EM iterations of the given phase (captured once per device state; the
        iteration index is a device counter, so the copies are identical launch sequences)."""
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (bool(nonrigid), int(unroll))
        if key not in graphs:
            g = torch.cuda.CUDAGraph()
            n0 = self._lib.spb_launch_count()
            with torch.cuda.graph(g):
                for _ in range(unroll):
                    check(self._lib.spb_em_iteration_ex(C.byref(self._params), -1, 1 if nonrigid else 0,
                                                        _capi.current_stream_ptr()), "spb_em_iteration_ex(capture)")
            graphs[key] = (g, int(self._lib.spb_launch_count() - n0))
        return graphs[key]

    @torch.no_grad()
    def run(self):
        """morpho_class.py:242-313. Returns P [N_A, N_B | batch] (numpy) or None when ``materialize_P=False``."""
        if not getattr(self, "_prepared", False):
            self.prepare_device()
        with torch.cuda.device(self._dev):
            if self.iter_key_added is not None:
                self._state["hist"] = torch.empty((max(self.max_iter, 1), 3, self.ldx), dtype=torch.float32, device=self._dev)
                self._state["hist_sigma2"] = torch.zeros((max(self.max_iter, 1),), dtype=torch.float64, device=self._dev)
            self.run_em()
            self._finish()
        return self.P

    # ------------------------------------------------------------------------------------------------------------------
    # closing similarity + output wrapping (morpho_class.py:296-313, 1437-1528)
    # ------------------------------------------------------------------------------------------------------------------
    def _finish(self):
        lib, p, s = self._lib, self._params, self._state
        st = _capi.current_stream_ptr()
        dt = self._np_dtype
        D, NA = self.D, self.NA
        last_iter = max(self.max_iter - 1, 0)
        if self.sigma2_end is not None:
            sc = self._read_scalars()
            sc.sigma2 = float(self.sigma2_end)
            s["sc"].copy_(torch.from_numpy(np.frombuffer(bytes(sc), dtype=np.uint8).copy()))
            check(lib.spb_row_update(C.byref(p), st), "spb_row_update")
        if self.max_iter == 0:
            self._estep_only(0, st)
            check(lib.spb_rigid_moments(C.byref(p), st), "spb_rigid_moments")
        if self.return_mapping and self.SVI_mode:
            # full (non-SVI) posterior with the final parameters (morpho_class.py:300-302)
            self.SVI_mode = False
            p.svi, p.NBb = 0, self.NB
            p.seg1 = p.seg2 = self._choose_segments(self.ldx // _capi.ROW_TILE, self.NB)
            self._NBb = self.NB
            self._estep_only(last_iter, st)
            # scalar Sp's must become the un-averaged sums (morpho_class.py:1183-1185)
            sc = self._read_scalars()
            sc.Sp_spatial, sc.Sp_sigma2, sc.Sp = sc.sums[0], sc.sums[1], sc.sums[2]
            s["sc"].copy_(torch.from_numpy(np.frombuffer(bytes(sc), dtype=np.uint8).copy()))
            s["moments"].zero_()
            check(lib.spb_rigid_moments(C.byref(p), st), "spb_rigid_moments")
        check(lib.spb_optimal_rigid(C.byref(p), ptr(s["optimal"]), st), "spb_optimal_rigid")
        opt = s["optimal"].cpu().numpy()
        sc = self._read_scalars()
        R3 = np.array(list(sc.R), dtype=np.float64).reshape(3, 3)
        self.R = R3[:D, :D].astype(dt)
        self.t = np.array(list(sc.t), dtype=np.float64)[None, :D].astype(dt)
        self.optimal_R = opt[:9].reshape(3, 3)[:D, :D].astype(dt)
        self.optimal_t = opt[9 : 9 + D].astype(dt)
        self.sigma2 = np.asarray(sc.sigma2, dtype=dt)
        self.gamma = np.asarray(sc.gamma, dtype=dt)
        self.sigma2_variance = np.asarray(sc.sigma2_variance, dtype=dt)
        self.Sp, self.Sp_spatial, self.Sp_sigma2 = sc.Sp, sc.Sp_spatial, sc.Sp_sigma2
        self.nonrigid_flag = self.max_iter - 1 > self.nonrigid_start_iter

        def rows(name):  # [3, ldx] SoA (processing order) -> [NA, D] in the caller's row order
            t = s[name][:D, :NA].T.contiguous()
            _count_d2h(t)
            return self._unsorted(t.cpu().numpy().astype(dt))

        def vec(name):
            _count_d2h(s[name][:NA])
            return self._unsorted(s[name][:NA].cpu().numpy().astype(dt))

        self.XAHat, self.RnA, self.VnA = rows("XAHat"), rows("RnA"), rows("VnA")
        self.optimal_RnA = (self.coordsA.astype(np.float64) @ self.optimal_R.astype(np.float64).T + self.optimal_t).astype(dt)
        self.K_NA = vec("K_NA")
        self.K_NB = s["K_NB"][: self._NBb].cpu().numpy().astype(dt)
        if self.column_shard is not None:  # every rank holds the column sums of its own block of fixed cells
            import torch.distributed as dist

            if dist.is_initialized() and dist.get_world_size() > 1:
                world = dist.get_world_size()
                width = (self.NB + world - 1) // world + 1
                mine = torch.zeros((width,), dtype=torch.float32, device=self._dev)
                mine[: self._NBb] = s["K_NB"][: self._NBb]
                parts = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                sizes = [(self.NB * (r + 1)) // world - (self.NB * r) // world for r in range(world)]
                self.K_NB = np.concatenate([q[:n].cpu().numpy() for q, n in zip(parts, sizes)]).astype(dt)
        self.K_NA_spatial = vec("K_NA_spatial")
        self.K_NA_sigma2 = vec("K_NA_sigma2")
        self.alpha = vec("alpha")
        self.SigmaDiag = vec("SigmaDiag")
        if self.nonrigid_flag:
            self.Coff = s["Coff"][:, :D].cpu().numpy().astype(dt)
            self.SigmaInv = s["SigmaInv"].cpu().numpy().astype(dt)
        else:
            self.Coff = np.zeros(self.K, dtype=dt)  # the reference's initial value (morpho_class.py:733)
        self.trace = s["trace_buf"].cpu().numpy()
        if (self.materialize_P or self.compute_mapping) and getattr(self, "_P_dev", None) is None:
            self._capture_P(last_iter, st)  # max_iter == 0 or the return_mapping E-step above
        if self.compute_mapping:
            from .mapping import ArgmaxPi

            ra, rv = ArgmaxPi.decode(self._rowbest.cpu().numpy().view(np.uint64))
            ca, cv = ArgmaxPi.decode(self._colbest.cpu().numpy().view(np.uint64))
            if self._perm is not None:  # device rows are in processing order
                ra, rv, ca = self._unsorted(ra), self._unsorted(rv), self._perm[ca]
            self.mapping = ArgmaxPi((NA, self._NBb), ra, rv.astype(dt), ca, cv.astype(dt))
            self._rowbest = self._colbest = None
        if self.materialize_P:
            if self.sparse_calculation_mode:
                self.P = self._sparse_P_to_coo(dt)
                self._P_rows = self._P_vals = None
            else:
                _count_d2h(self._P_dev)
                self.P = self._unsorted(self._P_dev.cpu().numpy().astype(dt))
        else:
            self.P = None
        self._P_dev = None
        if self.iter_key_added is not None:
            hist_d = s["hist"][:, :D, :NA].permute(0, 2, 1).contiguous()
            _count_d2h(hist_d)
            hist = hist_d.cpu().numpy().astype(dt)
            del hist_d
            if self._perm is not None:
                un = np.empty_like(hist)
                un[:, self._perm, :] = hist
                hist = un
            sig = s["hist_sigma2"].cpu().numpy()
            self.iter_added = {self.key_added: {}, "sigma2": {}}
            for it in range(self.max_iter):
                xa = hist[it]
                if self.normalize_c:
                    xa = xa * self.normalize_scales[1] + self.normalize_means[1]
                self.iter_added[self.key_added][it] = xa
                self.iter_added["sigma2"][it] = np.asarray(sig[it], dtype=dt)
        self._wrap_output()

    def _wrap_output(self):
        if self.normalize_c:
            sc1, m1 = self.normalize_scales[1], self.normalize_means[1]
            self.XAHat = self.XAHat * sc1 + m1
            self.RnA = self.RnA * sc1 + m1
            self.optimal_RnA = self.optimal_RnA * sc1 + m1
        if self.vecfld_key_added is not None:
            norm_dict = {
                "mean_transformed": self.normalize_means[0],
                "mean_fixed": self.normalize_means[1],
                "scale": self.normalize_scales[0],
                "scale_transformed": self.normalize_scales[0],
                "scale_fixed": self.normalize_scales[1],
            } if self.normalize_c else None
            self.vecfld = {
                "R": self.R,
                "t": self.t,
                "optimal_R": self.optimal_R,
                "optimal_t": self.optimal_t,
                "init_R": self.init_R if self.nn_init else np.eye(self.D),
                "init_t": self.init_t if self.nn_init else np.zeros(self.D),
                "beta": self.beta,
                "Coff": self.Coff,
                "inducing_variables": self.inducing_variables,
                "normalize_scales": self.normalize_scales if self.normalize_c else None,
                "normalize_means": self.normalize_means if self.normalize_c else None,
                "normalize_c": self.normalize_c,
                "dissimilarity": self.dissimilarity,
                "sigma2": self.sigma2,
                "gamma": self.gamma,
                "NA": self.NA,
                "sigma2_variance": self.sigma2_variance,
                "method": "Spateo",
                "norm_dict": norm_dict,
                "kernel_type": self.kernel_type,
            }

    # ------------------------------------------------------------------------------------------------------------------
    # test / debugging access to the device state
    # ------------------------------------------------------------------------------------------------------------------
    def device_vector(self, name: str) -> np.ndarray:
        t = self._state[name]
        return t.cpu().numpy()
