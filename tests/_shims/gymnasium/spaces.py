"""Stand-in for `gymnasium.spaces` (see package docstring)."""
import numpy as np


class Space:
    shape = None
    dtype = None


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(int(s) for s in shape)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class MultiBinary(Space):
    def __init__(self, n):
        self.n = n
        self.shape = (int(n),)
        self.dtype = np.dtype(np.int8)


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)


def flatdim(space):
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    if isinstance(space, Discrete):
        return space.n
    return int(np.prod(space.shape))


def flatten_space(space):
    if isinstance(space, Tuple):
        parts = [flatten_space(s) for s in space.spaces]
        return Box(
            low=np.concatenate([p.low for p in parts]),
            high=np.concatenate([p.high for p in parts]),
            dtype=np.result_type(*[p.dtype for p in parts]),
        )
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, MultiBinary):
        return Box(low=0, high=1, shape=(flatdim(space),), dtype=space.dtype)
    raise NotImplementedError(type(space))


def flatten(space, x):
    if isinstance(space, Tuple):
        return np.concatenate([np.array(flatten(s, xi)) for xi, s in zip(x, space.spaces)])
    if isinstance(space, (Box, MultiBinary)):
        return np.asarray(x, dtype=space.dtype).flatten()
    raise NotImplementedError(type(space))


def unflatten(space, x):
    if isinstance(space, Tuple):
        dims = np.asarray([flatdim(s) for s in space.spaces], dtype=np.int_)
        parts = np.split(x, np.cumsum(dims[:-1]))
        return tuple(unflatten(s, p) for p, s in zip(parts, space.spaces))
    if isinstance(space, (Box, MultiBinary)):
        return np.asarray(x, dtype=space.dtype).reshape(space.shape)
    raise NotImplementedError(type(space))
