"""CPU: the golden fixtures and the oracle pin are produced by running the reference on the `gymnasium` stand-in in
tests/_shims (gymnasium is not installed in the build container).  Where the REAL package is installed, this test asserts that
the stand-in and the real package agree on everything the reference uses -- `flatten_space`, `flatten`, `unflatten`, `flatdim`,
dtypes and shapes -- for the observation spaces of every canonical configuration (advisor finding of round 1).  Skipped when
gymnasium is absent."""
import importlib
import os
import sys

import numpy as np
import pytest

SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")


def _real_gymnasium():
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if k == "gymnasium" or k.startswith("gymnasium.")}
    try:
        sys.path = [p for p in sys.path if os.path.abspath(p) != SHIMS]
        for k in saved_mods:
            del sys.modules[k]
        try:
            mod = importlib.import_module("gymnasium")
            importlib.import_module("gymnasium.spaces")
        except ImportError:
            return None
        if os.path.abspath(os.path.dirname(mod.__file__)).startswith(SHIMS):
            return None
        return mod
    finally:
        sys.path = saved_path
        for k in [k for k in sys.modules if k == "gymnasium" or k.startswith("gymnasium.")]:
            if k not in saved_mods:
                sys.modules.pop(k, None)
        sys.modules.update(saved_mods)


def _shim():
    spec = importlib.util.spec_from_file_location("shim_gymnasium", os.path.join(SHIMS, "gymnasium", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(SHIMS, "gymnasium")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["shim_gymnasium"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("A,J,tagging", [(2, 0, False), (3, 5, True), (5, 5, False), (5, 0, False), (8, 8, True)])
def test_shim_matches_real_gymnasium(A, J, tagging):
    real = _real_gymnasium()
    if real is None:
        pytest.skip("gymnasium is not installed")
    shim = _shim()
    rng = np.random.default_rng(A * 100 + J)

    def spaces(g):  # base.py:211-228, tagging.py:42-60
        s = [g.spaces.Box(low=0, high=8, shape=(A, 2), dtype=int), g.spaces.MultiBinary(A)]
        if J > 0 or tagging:
            s += [g.spaces.Box(low=0, high=8, shape=(J, 2), dtype=int), g.spaces.MultiBinary(J)]
        if tagging:
            s += [g.spaces.MultiBinary(A), g.spaces.Box(low=0, high=A, shape=(A,), dtype=int), g.spaces.Box(low=0, high=50, shape=(1,), dtype=int)]
        return g.spaces.Tuple(s)

    rs, ss = spaces(real), spaces(shim)
    state = [rng.integers(0, 9, (A, 2)), rng.integers(0, 2, A).astype(bool)]
    if J > 0 or tagging:
        state += [rng.integers(0, 9, (J, 2)), rng.integers(0, 2, J).astype(bool)]
    if tagging:
        state += [rng.integers(0, 2, A).astype(bool), rng.integers(0, A, A), np.array([17])]
    state = tuple(state)
    fr, fs = real.spaces.flatten(rs, state), shim.spaces.flatten(ss, state)
    assert fr.dtype == fs.dtype and np.array_equal(fr, fs)
    assert real.spaces.flatten_space(rs).shape == shim.spaces.flatten_space(ss).shape
    assert real.spaces.flatten_space(rs).dtype == shim.spaces.flatten_space(ss).dtype
    ur, us = real.spaces.unflatten(rs, fr), shim.spaces.unflatten(ss, fs)
    assert len(ur) == len(us)
    for a, b in zip(ur, us):
        assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)
