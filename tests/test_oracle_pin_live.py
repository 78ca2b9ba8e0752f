"""Build-container only (skipped wherever /root/reference is absent, e.g. on the GPU box): the numpy oracle against the
UNMODIFIED reference executed live, two of the nine configurations of ``oracle/check_oracle_vs_reference.py`` — every output
identical to the last bit. The committed fixtures (tests/golden) carry the same pin to machines without the reference."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.ref_harness import reference_available  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("name, args, kw", [
    ("2d-full-f32", (300, 280, 30, 2, "float32", False, 100), {}),
    ("3d-svi-f32-sparse", (600, 560, 24, 3, "float32", True, 95), dict(sparse_calculation_mode=True, sparse_top_k=32)),
])
def test_oracle_is_bitwise_equal_to_the_reference(name, args, kw, capsys):
    from oracle.check_oracle_vs_reference import run_case

    assert run_case(name, *args, **kw) == 0.0
