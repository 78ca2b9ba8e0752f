"""Optimal-mapping helpers (spateo/alignment/utils.py:157-254) against the fixture generated from the unmodified reference
(tests/golden/make_golden_mapping.py): the oracle restatement and the host assembly logic on CPU, the device reductions and
the fused posterior-argmax kernels on the GPU."""

import numpy as np
import pytest

from oracle import mapping_oracle as mo_map


def _argmax_pi_numpy(pi):
    from spateo_release_b200.alignment.mapping import ArgmaxPi

    return ArgmaxPi(pi.shape, pi.argmax(1), pi.max(1), pi.argmax(0), pi.max(0))


@pytest.mark.parametrize("tag", ["2d", "3d"])
@pytest.mark.parametrize("keep_all", [False, True])
def test_mapping_oracle_and_host_assembly(golden, tag, keep_all):
    from spateo_release_b200.alignment.mapping import get_optimal_mapping_relationship

    g = golden("mapping")
    X, Y, pi = g[f"{tag}_X"], g[f"{tag}_Y"], g[f"{tag}_pi"]
    sfx = "_all" if keep_all else ""
    want = [g[f"{tag}_{k}{sfx}"] for k in ("xi", "xv", "yi", "yv")]
    got = mo_map.get_optimal_mapping_relationship(X, Y, pi, keep_all)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # product host logic on maxima computed with numpy (no device needed for an ArgmaxPi)
    got = get_optimal_mapping_relationship(X, Y, _argmax_pi_numpy(pi), keep_all)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b)
    gotT = get_optimal_mapping_relationship(Y, X, _argmax_pi_numpy(pi).T, keep_all)
    assert sorted(map(tuple, gotT[0][:, ::-1])) == sorted(map(tuple, want[2]))


def test_mapping_aligned_coords_from_argmax(golden):
    from spateo_release_b200.alignment.mapping import mapping_aligned_coords

    g = golden("mapping")
    X, Y, pi = g["2d_X"], g["2d_Y"], g["2d_pi"]
    mx, my = mapping_aligned_coords(X, Y, _argmax_pi_numpy(pi))
    for nm, mp in (("mx", mx), ("my", my)):
        for k, v in mp.items():
            assert np.allclose(v, g[f"2d_{nm}_{k}"]), (nm, k)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["2d", "3d"])
@pytest.mark.parametrize("keep_all", [False, True])
def test_mapping_dense_pi_on_device(golden, tag, keep_all):
    from spateo_release_b200.alignment.mapping import get_optimal_mapping_relationship, mapping_aligned_coords

    g = golden("mapping")
    X, Y, pi = g[f"{tag}_X"], g[f"{tag}_Y"], g[f"{tag}_pi"]
    sfx = "_all" if keep_all else ""
    got = get_optimal_mapping_relationship(X, Y, pi, keep_all)
    for a, k in zip(got, ("xi", "xv", "yi", "yv")):
        assert np.array_equal(a, g[f"{tag}_{k}{sfx}"]), k
    if not keep_all:
        mx, my = mapping_aligned_coords(X, Y, pi)
        for nm, mp in (("mx", mx), ("my", my)):
            for k, v in mp.items():
                assert np.allclose(v, g[f"{tag}_{nm}_{k}"]), (nm, k)


@pytest.mark.gpu
@pytest.mark.parametrize("svi", [False, True])
def test_fused_posterior_argmax_equals_dense_P(svi):
    """compute_mapping: the fused kernels' maxima are those of the materialised P of the same run."""
    import spateo_release_b200 as st
    from spateo_release_b200.synthetic import make_slice_pair

    A, B = make_slice_pair(2300, 2100, 40, dim=3, seed=4)
    np.random.seed(0)
    m = st.align.Morpho_pairwise(sampleA=B, sampleB=A, SVI_mode=svi, max_iter=60, nonrigid_start_iter=30, verbose=False,
                                 device="0", compute_mapping=True)
    P = m.run()
    mp = m.mapping
    assert mp.shape == P.shape
    assert np.array_equal(mp.row_val, P.max(1)) and np.array_equal(mp.col_val, P.max(0))
    nz = P.max(1) > 0
    assert np.array_equal(mp.row_arg[nz], P.argmax(1)[nz])
    assert np.array_equal(mp.col_arg[P.max(0) > 0], P.argmax(0)[P.max(0) > 0])
    Xa, Yb = m.optimal_RnA, np.asarray(A.obsm["spatial"])[: P.shape[1]] if not svi else np.asarray(A.obsm["spatial"])[m.batch_idx]
    a = st.align.get_optimal_mapping_relationship(Xa, Yb, mp)
    b = st.align.get_optimal_mapping_relationship(Xa, Yb, P)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    # without the dense posterior
    np.random.seed(0)
    m2 = st.align.Morpho_pairwise(sampleA=B, sampleB=A, SVI_mode=svi, max_iter=60, nonrigid_start_iter=30, verbose=False,
                                  device="0", compute_mapping=True, materialize_P=False)
    assert m2.run() is None
    assert np.array_equal(m2.mapping.row_arg[nz], mp.row_arg[nz]) and np.array_equal(m2.mapping.col_val, mp.col_val)
