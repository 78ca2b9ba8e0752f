"""Micro-benchmark: fused persistent E-step (one cooperative launch, second GT read from L2) against the three-kernel
path, at several panel sizes / L2-hint settings and at two EM states (dense early iteration, culled late iteration).
CUDA events on the launching stream, 100k x 100k pair by default."""
import argparse
import copy
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
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--panels", default="16,28,40")
ap.add_argument("--late-iter", type=int, default=150)
a = ap.parse_args()
dev = torch.device("cuda", 0)
A, B = bench.make_pair_on_device(a.cells, a.genes, 3, 0, dev)
lib = _capi.load_library()
stp = _capi.current_stream_ptr()


def build(panel_mb, fuse):
    os.environ["SPB_FUSE_PANEL_MB"] = str(panel_mb)
    np.random.seed(0)
    m = st.align.Morpho_pairwise(B, A, SVI_mode=False, max_iter=200, K=15, nn_init=False, verbose=False, device="0",
                                 materialize_P=False, fuse_estep=fuse)
    return m


def time_estep(m, it, reps):
    ts = []
    for rep in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        m._estep_only(it, stp)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            ts.append(e0.elapsed_time(e1))
    sc = m._read_scalars()
    return float(np.mean(ts)), float(np.min(ts)), (sc.sums[0], sc.sums[2], sc.sums[3]), sc.visited


m = build(28, False)
m.prepare()
GTkeep = m._GT
states = {}
m.run_em(n_iter=3)
torch.cuda.synchronize()
states["early(it3)"] = {k: v.clone() for k, v in m._state.items() if torch.is_tensor(v)}
m.run_em(n_iter=a.late_iter - 3, start=3)
torch.cuda.synchronize()
states[f"late(it{a.late_iter})"] = {k: v.clone() for k, v in m._state.items() if torch.is_tensor(v)}


def restore(mm, snap):
    for k, v in snap.items():
        if k in mm._state and torch.is_tensor(mm._state[k]) and mm._state[k].shape == v.shape:
            mm._state[k].copy_(v)


pairs = float(m.NA) * m.NB
for name, snap in states.items():
    restore(m, snap)
    mean, mn, sums, vis = time_estep(m, 5, a.reps)
    nrb = m.ldx // 1024
    frac = vis / (nrb * m.NB)
    print(f"[{name}] unfused: {mean:.3f} ms (min {mn:.3f})  visited {frac:.3f}  eff {8 * pairs * frac / mean / 1e6:.0f} GB/s  sums {sums}", flush=True)
    for hints in (0, 1):
        lib.spb_set_sweep_config(10 + hints)
        for pmb in [float(x) for x in a.panels.split(",")]:
            os.environ["SPB_FUSE_PANEL_MB"] = str(pmb)
            mf = copy.copy(m)  # shares the resident cost matrix / kernel matrices; gets its own EM state
            mf.fuse_estep = True
            with torch.cuda.device(dev):
                mf._allocate_state()
            restore(mf, snap)
            mean, mn, sums, vis = time_estep(mf, 5, a.reps)
            p = mf._params
            print(f"[{name}] fused panel {pmb:.0f} MB (W={p.fuse_W}, nseg={p.fuse_nseg}) hints={hints}: {mean:.3f} ms (min {mn:.3f})  "
                  f"eff {8 * pairs * frac / mean / 1e6:.0f} GB/s  sums {sums}", flush=True)
            del mf
            torch.cuda.empty_cache()
lib.spb_set_sweep_config(10)
