"""CPU: pin the C oracle against the UNMODIFIED reference run right here (needs the reference's sources: `/root/reference`
in the build container or the staged `baseline/_ref`; skipped otherwise -- the golden fixtures in tests/golden/ carry the
same pin to machines without the reference).  Same checks as `tools/check_oracle_vs_reference.py` at sizes that take
seconds: identical flat states, float64 reward bit patterns, dones, truncations, per-episode metrics, post-reset states,
role masks, sampled actions, running returns, and every featurizer's output."""
import importlib.util
import os

import pytest

from oracle import ref_harness as H
from tests.cases import CASES, EDGE_CASES

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference sources not available")

_spec = importlib.util.spec_from_file_location(
    "check_oracle_vs_reference", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                                              "check_oracle_vs_reference.py"))
pin = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(pin)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_equals_reference_on_canonical_cases(name):
    _, steps, episodes = pin.run_case((name, 6, 70, 31337, 40))
    assert steps == 6 * 70


@pytest.mark.parametrize("name", list(EDGE_CASES))
def test_oracle_equals_reference_on_edge_cases(name):
    pin.run_case((("edge", name), 4, 45, 5, 0))


@pytest.mark.parametrize("k", range(8))
def test_oracle_equals_reference_on_random_constructor_arguments(k):
    pin.run_case((("random", k), 4, 50, 900 + k, 3 * k))
