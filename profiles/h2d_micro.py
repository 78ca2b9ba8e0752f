"""Micro-benchmark (GPU box): ways of getting an 800 MB pageable numpy array onto the device — pageable copy, pin_memory +
async copy, cudaHostRegister in place, and a double-buffered staging pipeline through two reusable pinned buffers."""
import time

import numpy as np
import torch

n, G = 100000, 2000
x = np.random.default_rng(0).poisson(1.0, size=(n, G)).astype(np.float32)
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
rt = torch.cuda.cudart()


def timed(f, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y = f()
        torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
        del y
    return min(out), out


def pageable():
    return torch.from_numpy(x).to(dev)


def pin_then_copy():
    return torch.from_numpy(x).pin_memory().to(dev, non_blocking=True)


def register_in_place():
    t = torch.from_numpy(x)
    rc = rt.cudaHostRegister(t.data_ptr(), t.numel() * 4, 0)
    assert int(rc) == 0, rc
    y = torch.empty(t.shape, dtype=t.dtype, device=dev)
    y.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    rt.cudaHostUnregister(t.data_ptr())
    return y


stage = [torch.empty((8192, G), dtype=torch.float32).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event() for _ in range(2)]


def staged():
    t = torch.from_numpy(x)
    y = torch.empty(t.shape, dtype=t.dtype, device=dev)
    rows = stage[0].shape[0]
    for k, r0 in enumerate(range(0, n, rows)):
        b = k & 1
        r1 = min(n, r0 + rows)
        if k >= 2:
            ev[b].synchronize()
        stage[b][: r1 - r0].copy_(t[r0:r1])
        y[r0:r1].copy_(stage[b][: r1 - r0], non_blocking=True)
        ev[b].record()
    return y


for name, f in (("pageable .to()", pageable), ("pin_memory + async copy", pin_then_copy),
                ("cudaHostRegister in place", register_in_place), ("staged through 2 pinned buffers", staged)):
    best, all_ = timed(f)
    print(f"{name:34s} best {best * 1e3:8.1f} ms  ({x.nbytes / best / 1e9:6.1f} GB/s)  all {[round(a * 1e3) for a in all_]}")
