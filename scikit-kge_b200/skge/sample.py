"""Negative samplers on the device (reference: skge/sample.py).

``RandomModeSampler`` and ``LCWASampler`` keep the reference's constructor and
``sample(xys)`` contract (a list of ``((s, o, p), -1.0)``, subject-corrupted
first, skipping slots that exhausted ``ntries``); the draws happen in one
kernel (``skge_sample_corrupt``) against a device hash set of the training
triples.  The stream of random numbers is Philox, not numpy's MT19937: parity
with the reference is distributional.  ``RandomSampler`` / ``CorruptedSampler``
are dead code in the reference (their constructors raise) and are not provided.
"""
from collections import defaultdict as ddict

import numpy as np
import torch

from . import _ext, kernels


class Sampler(object):

    def __init__(self, n, modes, ntries=100):
        self.n = n
        self.modes = modes
        self.ntries = ntries
        self.seed = 42
        self._calls = 0
        self._set = None

    # -- device side -----------------------------------------------------------
    lcwa = False

    def device_ready(self):
        return hasattr(self, '_xs_arr')

    def ensure_device(self, xs=None):
        if self._set is None:
            a = self._xs_arr
            t = torch.from_numpy(np.ascontiguousarray(a.T, dtype=np.int32)).to(_ext.device())
            self._set = kernels.TripleSet(t[0].contiguous(), t[1].contiguous(), t[2].contiguous(),
                                          self.sz[0], self.sz[2], lcwa=self.lcwa)
        return self._set

    def train_size(self):
        return self._xs_arr.shape[0]

    def _modes_mask(self):
        modes = sorted(set(int(m) for m in self.modes))
        if list(self.modes) != modes or any(m not in (0, 1, 2) for m in modes):
            raise ValueError('modes must be an ascending subset of [0, 1, 2], got %r' % (self.modes,))
        return sum(1 << m for m in modes)

    def device_sample(self, batch_idx, B, call_no, src=None, offset_dev=None, outs=None, valid=None):
        """(pos, neg, valid) for B positives: indices ``batch_idx`` into the
        training arrays, or the explicit ``src`` = (s, o, p) tensors.  Calls get
        disjoint Philox counter ranges: by ``call_no`` (host) or by ``offset_dev``, an
        int64 CUDA scalar the caller advances (graph-captured steps)."""
        ts = self.ensure_device()
        return ts.sample(batch_idx, B, self.n, self._modes_mask(), self.ntries, self.seed,
                         (call_no << 40), src=src, offset_dev=offset_dev, outs=outs, valid=valid)

    # -- reference API ------------------------------------------------------------
    def sample(self, xys):
        if len(xys) == 0:
            return []
        a = np.array([x for x, _ in xys], dtype=np.int32).reshape(-1, 3)
        t = torch.from_numpy(np.ascontiguousarray(a.T)).to(_ext.device())
        _, neg, valid = self.device_sample(None, a.shape[0], self._calls,
                                           src=(t[0].contiguous(), t[1].contiguous(), t[2].contiguous()))
        self._calls += 1
        neg = torch.stack(neg, 1).cpu().numpy()
        valid = valid.cpu().numpy().astype(bool)
        return [((int(s), int(o), int(p)), -1.0) for (s, o, p), v in zip(neg, valid) if v]


class RandomModeSampler(Sampler):
    """Sample negative triples by corrupting one slot uniformly at random,
    rejecting training triples (skge/sample.py:28-46)."""

    def __init__(self, n, modes, xs, sz):
        super(RandomModeSampler, self).__init__(n, modes)
        self._xs_arr = np.asarray(xs, dtype=np.int64).reshape(-1, 3)
        self.sz = sz

    @property
    def xs(self):
        return set(map(tuple, self._xs_arr.tolist()))


class LCWASampler(RandomModeSampler):
    """Local-closed-world sampling: the corrupted triple's (s, p) must occur in
    the training set (skge/sample.py:91-110)."""

    lcwa = True

    @property
    def counts(self):
        c = ddict(int)
        for s, o, p in self._xs_arr.tolist():
            c[(s, p)] += 1
        return c


def type_index(xs):
    """skge/sample.py:113-120 (used only by the reference's dead CorruptedSampler)."""
    index = ddict(lambda: {0: set(), 1: set()})
    for i, j, k in xs:
        index[k][0].add(i)
        index[k][1].add(j)
    return {k: {0: list(v[0]), 1: list(v[1])} for k, v in index.items()}
