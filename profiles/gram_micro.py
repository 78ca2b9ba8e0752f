"""Micro-run of the tcgen05 K^T P K contraction for ncu (SparseVFC shape by default: N = 1e6, K = 500)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from spateo_release_b200 import _capi  # noqa: E402
from spateo_release_b200._capi import check, ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1000000)
ap.add_argument("--K", type=int, default=500)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
lib = _capi.load_library()
dev = torch.device("cuda", 0)
N, K = a.N, a.K
ldn = ((N + 1023) // 1024) * 1024
g = torch.Generator(device=dev)
g.manual_seed(0)
UT = torch.rand((K, ldn), generator=g, device=dev, dtype=torch.float32)
w = torch.rand((ldn,), generator=g, device=dev, dtype=torch.float32)
X3 = torch.randn((3, ldn), generator=g, device=dev, dtype=torch.float32)
st = _capi.current_stream_ptr()
hi, lo = torch.empty_like(UT), torch.empty_like(UT)
mean = torch.empty((K,), dtype=torch.float32, device=dev)
check(lib.spb_gram_center(ptr(UT), ldn, N, K, ptr(mean), ptr(hi), ptr(lo), st), "center")
Bhi = torch.empty((K + 4, ldn), dtype=torch.float32, device=dev)
Blo = torch.empty_like(Bhi)
sums = torch.empty((4,), dtype=torch.float64, device=dev)
need = C.c_int64(0)
check(lib.spb_gram_tc_scratch_floats(K, 3, N, C.byref(need)), "plan")
scratch = torch.empty((need.value,), dtype=torch.float32, device=dev)
G = torch.empty((K, K), dtype=torch.float64, device=dev)
R = torch.empty((K, 3), dtype=torch.float64, device=dev)
ts = []
for rep in range(a.reps):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    check(lib.spb_gram_prepare(ptr(UT), ldn, N, K, ptr(mean), ptr(w), ptr(X3), ldn, 3, ptr(Bhi), ptr(Blo), ptr(sums), st), "prepare")
    e[1].record()
    check(lib.spb_gram_tc(ptr(hi), ptr(lo), ptr(Bhi), ptr(Blo), ldn, N, K, 3, ptr(mean), ptr(sums), ptr(scratch), scratch.numel(),
                          ptr(G), ptr(R), st), "gram_tc")
    e[2].record()
    torch.cuda.synchronize()
    ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
print(f"N={N} K={K}: prepare {min(t[0] for t in ts):.3f} ms, gram_tc {min(t[1] for t in ts):.3f} ms "
      f"({2.0 * N * K * (K + 3) / (min(t[1] for t in ts) * 1e-3) / 1e12:.1f} algorithmic TFLOP/s)")
