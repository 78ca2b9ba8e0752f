"""Small driver for ncu captures: one full-size slice pair, a few EM iterations (run under `ncu -k regex:estep_sweep`)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import spateo_release_b200 as st  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--genes", type=int, default=2000)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--start", type=int, default=0)
ap.add_argument("--nn-init", type=int, default=0)
ap.add_argument("--svi", type=int, default=0)
ap.add_argument("--K", type=int, default=15)
ap.add_argument("--graph", type=int, default=0, help="1: replay iterations from CUDA graphs (the product default)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
A, B = bench.make_pair_on_device(a.cells, a.genes, 3, 0, dev)
np.random.seed(0)
m = st.align.Morpho_pairwise(B, A, SVI_mode=bool(a.svi), max_iter=200, K=a.K, nn_init=bool(a.nn_init), verbose=False, device="0",
                             materialize_P=False)
m.prepare()
m.use_cuda_graph = bool(a.graph)
m.run_em(n_iter=a.iters, start=a.start)
torch.cuda.synchronize()
print("done", m._read_scalars().sigma2)
