"""Micro-benchmark of the two E-step sweep kernels (CUDA events, 100k x 100k pair): pipeline shapes 0..2 and the
diagnostic modes of sweep 2 (cfg 16 = bulk-copy pipeline alone, cfg 32 = arithmetic alone on a never-refilled ring)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import spateo_release_b200 as st  # noqa: E402
from spateo_release_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--genes", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--cfgs", default="0,16,32", help="pipeline shape + 16 * diagnostic mode (1 = sweep 2 streams only, 2 = sweep 2 arithmetic only)")
ap.add_argument("--warm-iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
A, B = bench.make_pair_on_device(a.cells, a.genes, 3, 0, dev)
np.random.seed(0)
m = st.align.Morpho_pairwise(B, A, SVI_mode=False, max_iter=200, K=15, nn_init=False, verbose=False, device="0",
                             materialize_P=False)
m.prepare()
lib = _capi.load_library()
m.run_em(n_iter=a.warm_iters)
torch.cuda.synchronize()
p = m._params
stp = _capi.current_stream_ptr()
for cfg in [int(c) for c in a.cfgs.split(",")]:
    _capi.check(lib.spb_set_sweep_config(cfg), "cfg")
    t1 = t2 = 0.0
    for rep in range(a.reps + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        lib.spb_iter_begin(C.byref(p), 3, stp)
        e[0].record(); lib.spb_estep_sweep1(C.byref(p), 3, stp); e[1].record()
        lib.spb_col_finalize(C.byref(p), stp)
        e[2].record(); lib.spb_estep_sweep2(C.byref(p), 3, stp); e[3].record()
        lib.spb_row_finalize(C.byref(p), stp)
        torch.cuda.synchronize()
        if rep >= 2:
            t1 += e[0].elapsed_time(e[1]); t2 += e[2].elapsed_time(e[3])
    sc = m._read_scalars()
    gb = 4.0 * m.NA * m.NB / 1e9
    print(f"cfg {cfg}: sweep1 {t1 / a.reps:.3f} ms ({gb / (t1 / a.reps) :.0f} GB/s)  sweep2 {t2 / a.reps:.3f} ms ({gb / (t2 / a.reps):.0f} GB/s)  "
          f"sums {sc.sums[0]:.6e} {sc.sums[2]:.6e} {sc.sums[3]:.6e}", flush=True)
