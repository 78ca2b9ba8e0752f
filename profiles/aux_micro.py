"""Timings of the kernels outside the dense EM loop (CUDA events, 1 x B200): sparse_calculation_mode E-step (exact
per-column top-k select), fused posterior arg-maxima, COO emission, field differential geometry, device voxelisation."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import spateo_release_b200 as st  # noqa: E402
from spateo_release_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--genes", type=int, default=256)
ap.add_argument("--topk", type=int, default=1024)
ap.add_argument("--late-iter", type=int, default=150)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _capi.load_library()
stp = _capi.current_stream_ptr()


def timed(fn, reps=5, warm=2):
    ts = []
    for r in range(warm + reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


A, B = bench.make_pair_on_device(a.cells, a.genes, 3, 0, dev)
out = {}
for sparse in (False, True):
    np.random.seed(0)
    m = st.align.Morpho_pairwise(B, A, SVI_mode=False, max_iter=200, K=15, nn_init=False, verbose=False, device="0",
                                 materialize_P=sparse, compute_mapping=True, sparse_calculation_mode=sparse,
                                 sparse_top_k=a.topk)
    m.prepare()
    tag = f"sparse(k={a.topk})" if sparse else "dense"
    for name, upto in (("early(it3)", 3), (f"late(it{a.late_iter})", a.late_iter)):
        start = 0 if upto == 3 else 3
        m.run_em(n_iter=upto - start, start=start)
        torch.cuda.synchronize()
        ms = timed(lambda: m._estep_only(upto, stp))
        sc = m._read_scalars()
        print(f"[{tag} {name}] E-step {ms:.3f} ms  Sp {sc.sums[2]:.6e}  Sp_spatial {sc.sums[0]:.6e}", flush=True)
        out[(tag, name)] = ms
        if sparse:
            p = m._params
            ms_sel = timed(lambda: _capi.check(lib.spb_estep_col_select(C.byref(p), upto, stp), "sel"))
            rows = torch.zeros((m.NB, a.topk), dtype=torch.int32, device=dev)
            vals = torch.zeros((m.NB, a.topk), dtype=torch.float32, device=dev)
            ms_emit = timed(lambda: _capi.check(lib.spb_sparse_P_emit(C.byref(p), upto, _capi.ptr(rows), _capi.ptr(vals), stp), "emit"))
            print(f"    col_select alone {ms_sel:.3f} ms, COO emission ({m.NB} x {a.topk} entries) {ms_emit:.3f} ms", flush=True)
            del rows, vals
        else:
            p = m._params
            rb = torch.zeros((m.NA,), dtype=torch.int64, device=dev)
            cb = torch.zeros((m.NB,), dtype=torch.int64, device=dev)
            ms_arg = timed(lambda: _capi.check(lib.spb_posterior_argmax(C.byref(p), upto, _capi.ptr(rb), _capi.ptr(cb), stp), "argmax"))
            print(f"    posterior row+column arg-maxima {ms_arg:.3f} ms", flush=True)
    if sparse:  # size-independent properties of a full sparse run state
        kna = m._state["K_NA"][: m.NA].double().sum().item()
        knb = m._state["K_NB"][: m.NB].double().sum().item()
        print(f"    sum K_NA {kna:.6e}  sum K_NB {knb:.6e}  rel diff {abs(kna - knb) / kna:.2e}")
    del m
    torch.cuda.empty_cache()

# field geometry: 1M query points, K = 15 and K = 500 control points
from spateo_release_b200.tdr.morphofield_dg import field_geometry  # noqa: E402

rng = np.random.default_rng(0)
for K in (15, 500):
    vf = dict(norm_dict=dict(mean_transformed=np.zeros(3), scale_transformed=np.float64(30.0), mean_fixed=np.zeros(3),
                             scale_fixed=np.float64(30.0)),
              kernel_type="euc", inducing_variables=rng.normal(size=(K, 3)), beta=0.5, Coff=rng.normal(size=(K, 3)) * 0.05,
              R=np.eye(3), t=np.zeros((1, 3)))
    X = rng.normal(size=(1000000, 3)) * 30
    t0 = time.perf_counter()
    g = field_geometry(X, vf, want=("V", "J", "acc", "curv", "curl", "torsion", "div", "det"))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    g = field_geometry(X, vf, want=("V", "J", "acc", "curv", "curl", "torsion", "div", "det"))
    torch.cuda.synchronize()
    print(f"[field geometry, 1M points, K={K}] {1e3 * (time.perf_counter() - t1):.1f} ms per call incl. H2D/D2H of all outputs "
          f"(first call {1e3 * (t1 - t0):.1f} ms)", flush=True)
