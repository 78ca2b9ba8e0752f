"""The reference's own unit tests for this path (tests/alignment/test_utils.py: ``check_rep_layer``, the only test module the
reference has for morpho-align) restated against the host mirror: same scenarios, same outcomes (True / ValueError)."""

import numpy as np
import pandas as pd
import pytest

from spateo_release_b200.alignment import utils as U
from spateo_release_b200.anndata_lite import AnnDataLite


def _sample(n, g=10, d=3, seed=0):
    rng = np.random.default_rng(seed)
    obs = pd.DataFrame({"label": pd.Categorical(rng.choice(["A", "B", "C"], n)), "scalar": rng.normal(size=n)})
    return AnnDataLite(rng.normal(size=(n, g)), obs=obs, layers={"layer": rng.normal(size=(n, g))},
                       obsm={"rep": rng.normal(size=(n, d))})


@pytest.fixture(scope="module")
def samples():
    return [_sample(100, seed=1), _sample(200, seed=2)]


@pytest.mark.parametrize("layers, fields", [
    (["layer"], ["layer"]), (["X"], ["layer"]), (["rep"], ["obsm"]), (["label"], ["obs"]),
    (["X", "rep", "layer", "label"], ["layer", "obsm", "layer", "obs"]),
])
def test_valid_representations(samples, layers, fields):
    assert U.check_rep_layer(samples, rep_layer=layers, rep_field=fields) is True


@pytest.mark.parametrize("layers, fields", [
    (["invalid_layer"], ["layer"]), (["invalid_obsm"], ["obsm"]), (["invalid_obs"], ["obs"]),
    (["scalar"], ["obs"]),          # present but not categorical
    (["obs3"], ["obs"]), (["layer1"], ["invalid"]),
])
def test_invalid_representations(samples, layers, fields):
    with pytest.raises(ValueError):
        U.check_rep_layer(samples, rep_layer=layers, rep_field=fields)
