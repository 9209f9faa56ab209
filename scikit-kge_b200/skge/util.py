"""Host-side helpers with the reference's names (skge/util.py).

``ccorr`` / ``cconv`` / ``grad_sum_matrix`` are kept for callers that use them
on numpy arrays (the reference's scripts import them); the training and ranking
kernels never go through them -- their device counterparts live in
csrc/hole_math.cuh and csrc/segment.cu.
"""
import collections.abc
import functools

import numpy as np


def cconv(a, b):
    """Circular convolution c_k = sum_i a_i b_{(k-i) mod d}  (skge/util.py:8-27)."""
    return np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)).real


def ccorr(a, b):
    """Circular correlation c_k = sum_i a_i b_{(i+k) mod d}  (skge/util.py:30-50)."""
    return np.fft.ifft(np.conj(np.fft.fft(a)) * np.fft.fft(b)).real


def grad_sum_matrix(idx):
    """(sorted unique ids, selector matrix U x len(idx), occurrences per id)
    -- skge/util.py:53-101.  The selector is returned as a dense-free CSR
    triple wrapped in a tiny object with ``dot`` so ``Sm.dot(G) / n`` works."""
    uidx, iinv = np.unique(np.asarray(idx), return_inverse=True)
    n = np.bincount(iinv, minlength=len(uidx)).astype(np.float64)[:, np.newaxis]

    class _Selector(object):
        shape = (len(uidx), len(iinv))

        @staticmethod
        def dot(G):
            out = np.zeros((len(uidx),) + G.shape[1:], dtype=np.result_type(G, np.float64))
            np.add.at(out, iinv, G)
            return out

    return uidx, _Selector(), n


def unzip_triples(xys, with_ys=False):
    """[((s, o, p), y), ...] -> ss, ps, os[, ys]  (note the s, p, o return order;
    skge/util.py:104-110)."""
    xs, ys = list(zip(*xys))
    ss, os, ps = list(zip(*xs))
    if with_ys:
        return np.array(ss), np.array(ps), np.array(os), np.array(ys)
    return np.array(ss), np.array(ps), np.array(os)


class memoized(object):
    """Cache decorator with the reference's semantics (skge/util.py:134-164)."""

    def __init__(self, func):
        self.func = func
        self.cache = {}

    def __call__(self, *args):
        if not isinstance(args, collections.abc.Hashable):
            return self.func(*args)
        if args not in self.cache:
            self.cache[args] = self.func(*args)
        return self.cache[args]

    def __repr__(self):
        return self.func.__doc__

    def __get__(self, obj, objtype):
        return functools.partial(self.__call__, obj)
