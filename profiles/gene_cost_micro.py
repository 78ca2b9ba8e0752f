"""Times the expression-cost GEMM (100k x 100k x 2000, KL + gauss epilogue) and checks it against fp64 on a sub-block."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from spateo_release_b200 import _capi  # noqa: E402
from spateo_release_b200.alignment.morpho_class import GeneCostBuilder  # noqa: E402

n, G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 2000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
A = torch.poisson(torch.rand((n, G), device=dev, generator=g) * 2, generator=g)
B = torch.poisson(torch.rand((n, G), device=dev, generator=g) * 2, generator=g)
gc = GeneCostBuilder(_capi.load_library(), dev)
opA, rtA = gc.prepare(A, "kl", fixed=False)
opB, rtB = gc.prepare(B, "kl", fixed=True, centre=gc.centre_of(opA, A.shape[1]))
ldx = (n + 1023) // 1024 * 1024
GT = torch.empty((n, ldx), dtype=torch.float32, device=dev)
Xa = A[:512].double() + 0.01; Xa = Xa / Xa.sum(1, keepdim=True)
Yb = B[:512].double() + 0.01; Yb = Yb / Yb.sum(1, keepdim=True)
e = (Xa * torch.log(Xa + 1e-8)).sum(1, keepdim=True) - Xa @ torch.log(Yb + 1e-8).T
want = torch.exp(-e / 0.2).T
for backend in ("simt", "tensor"):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gc.cost(opA, rtA, opB, rtB, n, n, G, "kl", "gauss", 0.1, False, GT, ldx, backend=backend); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"[{backend}] gene_cost {n}x{n}x{G}: {ms:.1f} ms  {2.0 * n * n * G / ms / 1e9:.1f} TFLOP/s (incl. operand split)", flush=True)
    got = GT[:512, :512].double()
    err_e = (-0.2 * torch.log(got) - e.T)
    print(f"[{backend}] max rel err of g vs fp64: {float(((got - want).abs() / want).max()):.3e}  e: max abs {float(err_e.abs().max()):.3e} "
          f"rms {float(err_e.pow(2).mean().sqrt()):.3e} mean(bias) {float(err_e.mean()):.3e}", flush=True)
    # far corner of the matrix too (last tile rows/cols)
    print(f"[{backend}] checksum(first 4096 rows) {float(GT[:4096, :n].double().sum()):.10e}  corner {float(GT[n-1, n-1]):.6e}", flush=True)
sys.exit(0)
# accuracy on a 512 x 512 block against float64
Xa = A[:512].double() + 0.01; Xa = Xa / Xa.sum(1, keepdim=True)
Yb = B[:512].double() + 0.01; Yb = Yb / Yb.sum(1, keepdim=True)
e = (Xa * torch.log(Xa + 1e-8)).sum(1, keepdim=True) - Xa @ torch.log(Yb + 1e-8).T
want = torch.exp(-e / 0.2).T
got = GT[:512, :512].double()
print("max rel err of g vs fp64:", float(((got - want).abs() / want).max()), " max abs err of e:",
      float((-0.2 * torch.log(got) - e.T).abs().max()))
