"""GPU: column-sharded single pair. One process emulates W ranks on ONE GPU: W solver objects, each
holding a block of the fixed cells, stepped in lock-step with the cross-rank sum of the row statistics done by the test —
the same kernels (spb_row_fold / spb_row_stats_finalize) and the same host path the multi-GPU driver uses, so the sharded
arithmetic is checked against the unsharded solver without needing several devices."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 3])
def test_column_sharded_pair_matches_unsharded(world):
    import torch

    import spateo_release_b200 as st
    from spateo_release_b200.synthetic import make_slice_pair

    A, B = make_slice_pair(2600, 2300, 40, dim=3, seed=5, z_thickness=15.0, warp_amplitude=1.0)
    kw = dict(SVI_mode=False, max_iter=110, K=15, nn_init=True, verbose=False, device="0", materialize_P=False)
    np.random.seed(0)
    ref = st.align.Morpho_pairwise(sampleA=B, sampleB=A, **kw)
    ref.run()

    shards = []
    for r in range(world):
        np.random.seed(0)
        m = st.align.Morpho_pairwise(sampleA=B, sampleB=A, column_shard=(r, world, "nccl"), **kw)
        m.prepare_host()
        shards.append(m)
    for m in shards[1:]:  # the driver broadcasts rank 0's host initialisation
        for k in ("coordsA", "init_R", "init_t", "inlier_A", "inlier_B", "inlier_P", "sigma2", "_sigma2_init",
                  "probability_parameters", "samples_s"):
            setattr(m, k, getattr(shards[0], k))
    for m in shards:
        m.prepare_device()
    cols = [m._col_range() for m in shards]
    assert cols[0][0] == 0 and cols[-1][1] == ref.NB and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
    st_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    from spateo_release_b200._capi import check

    if world == 1:  # the solver's own sharded iteration path (fold + finish instead of the fused row_finalize)
        shards[0].run_em()
    for it in range(kw["max_iter"] if world > 1 else 0):
        nonrigid = it > ref.nonrigid_start_iter
        # E-step pieces up to the local fold on every shard, then the cross-shard sum, then the finish — what
        # _shard_row_statistics does around an all_reduce
        for m in shards:
            lib, p = m._lib, m._params
            check(lib.spb_iter_begin(C.byref(p), it, st_ptr), "iter_begin")
            check(lib.spb_gather_cols(C.byref(p), it, st_ptr), "gather")
            check(lib.spb_estep_col_lists(C.byref(p), st_ptr), "lists")
            check(lib.spb_estep_sweep1(C.byref(p), it, st_ptr), "s1")
            check(lib.spb_col_finalize(C.byref(p), st_ptr), "cf")
            check(lib.spb_estep_sweep2(C.byref(p), it, st_ptr), "s2")
            check(lib.spb_row_fold(C.byref(p), it & 1, st_ptr), "fold")
        n = 8 * shards[0].ldx
        views = [m._state["rowstat"][(it & 1) * n : ((it & 1) + 1) * n] for m in shards]
        total = torch.zeros_like(views[0])
        for v in views:  # rank order, like the peer-memory kernel
            total += v
        for m, v in zip(shards, views):
            v.copy_(total)
            lib, p = m._lib, m._params
            check(lib.spb_row_stats_finalize(C.byref(p), it & 1, st_ptr), "finalize")
            check(lib.spb_update_gamma_alpha(C.byref(p), st_ptr), "ga")
            if nonrigid:
                check(lib.spb_nonrigid_accumulate(C.byref(p), st_ptr), "acc")
                check(lib.spb_nonrigid_solve(C.byref(p), st_ptr), "solve")
                check(lib.spb_field_apply(C.byref(p), st_ptr), "apply")
            check(lib.spb_rigid_moments(C.byref(p), st_ptr), "mom")
            check(lib.spb_rigid_solve(C.byref(p), it, st_ptr), "rigid")
            check(lib.spb_row_update(C.byref(p), st_ptr), "rows")
    for m in shards:
        m._finish()
    scale = np.abs(ref.XAHat).max()
    for m in shards:
        assert np.abs(m.XAHat - ref.XAHat).max() < 2e-5 * scale
        assert np.abs(m.optimal_RnA - ref.optimal_RnA).max() < 2e-5 * scale
        assert abs(float(m.sigma2) - float(ref.sigma2)) < 1e-4 * float(ref.sigma2)
        assert np.abs(m.K_NA - ref.K_NA).max() < 1e-4 * np.abs(ref.K_NA).max()
    # replicas are bit-identical; every shard's K_NB block is the matching slice of the unsharded column sums
    for m in shards[1:]:
        assert np.array_equal(m.XAHat, shards[0].XAHat) and np.array_equal(m.K_NA, shards[0].K_NA)
    knb = np.concatenate([m._state["K_NB"][: c1 - c0].cpu().numpy() for m, (c0, c1) in zip(shards, cols)])
    assert np.abs(knb - ref.K_NB).max() < 1e-4 * np.abs(ref.K_NB).max()
