"""Minimal stand-in for the nine `gymnasium` symbols Sus-Net's env imports.

TEST INFRASTRUCTURE ONLY.  `gymnasium` is not installed in the build container
and there is no network; the reference (`/root/reference/src/environment/base.py:7-8`,
`tagging.py:3`) only uses `Env` as a base class, `spaces.{Discrete,Box,MultiBinary,
Tuple,flatten_space,flatten,unflatten}` and imports (never calls)
`envs.registration.register`.  Semantics follow gymnasium >= 0.26.
Used only by `tools/make_golden.py` / CPU tests that replay the reference here.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *args, **kwargs):
        pass
