"""Parameters, initialisers, post-hooks and sparse updaters on the device.

Mirrors skge/param.py of the reference (same names, arguments and error
behaviour); storage is an fp32 CUDA tensor and the update is one sm_100a kernel
(``skge_sparse_update``: SGD/AdaGrad row update fused with the row post-hook).
"""
import numpy as np
import torch

from . import _ext

# ---------------------------------------------------------------------------
# initialisers (reference: skge/param.py:11-54) -- host API kept for parity;
# Parameter itself draws on the device.
# ---------------------------------------------------------------------------


def init_unif(sz):
    """U(-1/sqrt(sz[0]), +1/sqrt(sz[0]))  (skge/param.py:11-19)."""
    bnd = 1 / np.sqrt(sz[0])
    return np.squeeze(np.random.uniform(low=-bnd, high=bnd, size=sz))


def init_nunif(sz):
    """Glorot-style U(+-sqrt(6)/sqrt(sz[0]+sz[1]))  (skge/param.py:23-50)."""
    bnd = np.sqrt(6) / np.sqrt(sz[0] + sz[1])
    init_nunif.counter += 1
    return np.squeeze(np.random.uniform(low=-bnd, high=bnd, size=sz))


init_nunif.counter = 0


def init_randn(sz):
    """Standard normal  (skge/param.py:53-54)."""
    return np.squeeze(np.random.randn(*sz))


_INITS = ('unif', 'nunif', 'randn')
_seed_state = {'seed': 42}   # the reference seeds numpy with 42 at import (skge/base.py:48)


def set_seed(seed):
    """Seed of the device generator used by Parameter initialisation."""
    _seed_state['seed'] = int(seed)
    _seed_state.pop('gen', None)


def _generator():
    if 'gen' not in _seed_state:
        g = torch.Generator(device=_ext.device())
        g.manual_seed(_seed_state['seed'])
        _seed_state['gen'] = g
    return _seed_state['gen']


def _device_init(shape, method):
    if method not in _INITS:
        raise ValueError('Unknown initialization (%s)' % method)
    if len(shape) != 2:
        raise ValueError('Shape must be of size 2')
    dev, g = _ext.device(), _generator()
    if method == 'randn':
        return torch.randn(shape, device=dev, dtype=torch.float32, generator=g)
    bnd = (1.0 / np.sqrt(shape[0])) if method == 'unif' else (np.sqrt(6) / np.sqrt(shape[0] + shape[1]))
    return (torch.rand(shape, device=dev, dtype=torch.float32, generator=g) * 2 - 1) * float(bnd)


# ---------------------------------------------------------------------------
# device array handed out by the gradient hooks
# ---------------------------------------------------------------------------


class DevArray(object):
    """A CUDA tensor that numpy can read (``np.asarray``) -- what the model
    hooks return for gradients and row ids."""

    __slots__ = ('t',)

    def __init__(self, t):
        self.t = t

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return tuple(self.t.shape)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return iter(np.asarray(self))

    def __getitem__(self, k):
        return np.asarray(self)[k]

    def __repr__(self):
        return 'DevArray(%r)' % (np.asarray(self),)


def _as_tensor(x, dtype):
    if isinstance(x, DevArray):
        x = x.t
    if isinstance(x, Parameter):
        x = x.data
    if isinstance(x, torch.Tensor):
        return x.to(device=_ext.device(), dtype=dtype).contiguous()
    npdt = {torch.float32: np.float32, torch.int32: np.int32}[dtype]
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=npdt)).to(_ext.device())


# ---------------------------------------------------------------------------
# Parameter (reference: an ndarray subclass, skge/param.py:57-105)
# ---------------------------------------------------------------------------


class Parameter(object):
    """Parameter(shape, init, name=None, post=None)

    ``data`` is the fp32 CUDA tensor.  A 3-D shape (M, d, d) is a stack of M
    independent 2-D initialisations (skge/param.py:62-64).  ``post`` is applied
    once at creation (skge/param.py:72-73).  Indexing returns host numpy copies;
    item assignment uploads (so ``m.E[...] = E0`` works as with the reference).
    """

    def __init__(self, *args, **kwargs):
        shape, method = args[0], args[1]
        self.name = kwargs.pop('name', None)
        self.post = kwargs.pop('post', None)
        value = kwargs.pop('value', None)
        if value is not None:
            self.data = _as_tensor(value, torch.float32).clone()
        elif len(shape) == 3:
            self.data = torch.stack([_device_init((shape[1], shape[2]), method) for _ in range(shape[0])])
        else:
            self.data = _device_init(tuple(shape), method)
        n = self.data.shape[0]
        dev = self.data.device
        # instrumentation counters of the fork (skge/param.py:83-86), bumped by the kernels
        self._update_counts = torch.zeros(n, dtype=torch.int32, device=dev)
        self._violations = torch.zeros(n, dtype=torch.int32, device=dev)
        self._neighbours = torch.zeros(n, dtype=torch.int32, device=dev)
        if self.post is not None and value is None:
            r = self.post(self)
            if r is not None and r is not self:
                self.data = _as_tensor(r, torch.float32)

    # -- ndarray-like surface -------------------------------------------------
    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return self.data.dim()

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def __len__(self):
        return self.data.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self.data.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    @staticmethod
    def _key(k):
        if isinstance(k, (DevArray, Parameter)):
            k = np.asarray(k)
        if isinstance(k, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(k)).to(_ext.device())
        if isinstance(k, list):
            return torch.as_tensor(k, device=_ext.device())
        if isinstance(k, tuple):
            return tuple(Parameter._key(x) for x in k)
        return k

    def __getitem__(self, k):
        return self.data[self._key(k)].detach().cpu().numpy()

    def __setitem__(self, k, v):
        if isinstance(v, (DevArray, Parameter)):
            v = np.asarray(v)
        v = torch.as_tensor(np.asarray(v, dtype=np.float32) if not isinstance(v, torch.Tensor) else v,
                            device=self.data.device, dtype=torch.float32)
        self.data[self._key(k)] = v

    def __iter__(self):
        return iter(np.asarray(self))

    def __repr__(self):
        return 'Parameter(name=%r, shape=%r, device=%s)' % (self.name, self.shape, self.data.device)

    # -- counters ---------------------------------------------------------------
    @property
    def updateCounts(self):
        return self._update_counts.cpu().tolist()

    @property
    def violations(self):
        return self._violations.cpu().tolist()

    @property
    def neighbours(self):
        return self._neighbours.cpu().tolist()

    # -- pickling: host float64 arrays, loadable without a GPU --------------------
    def __getstate__(self):
        return {'value': np.asarray(self, dtype=np.float64), 'name': self.name,
                'post': getattr(self.post, '__name__', None)}

    def __setstate__(self, st):
        self.name = st['name']
        self.post = {'normalize': normalize, 'normless1': normless1, None: None}[st['post']]
        self.data = _as_tensor(st['value'], torch.float32)
        n = self.data.shape[0]
        for a in ('_update_counts', '_violations', '_neighbours'):
            setattr(self, a, torch.zeros(n, dtype=torch.int32, device=self.data.device))

    @classmethod
    def from_reference(cls, arr):
        """Wrap a reference-style Parameter (an ndarray subclass instance with ``name`` /
        ``post`` attributes, as found in pickles written by the reference)."""
        post = getattr(arr, 'post', None)
        post = {'normalize': normalize, 'normless1': normless1}.get(getattr(post, '__name__', None))
        return cls(arr.shape, 'nunif', name=getattr(arr, 'name', None), post=post,
                   value=np.asarray(arr, dtype=np.float64))



class RefParameter(np.ndarray):
    """Stand-in for the reference's ``skge.param.Parameter`` (an ndarray subclass,
    skge/param.py:57-105), used only to read reference-written pickles."""

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, 'name', None)
        self.post = getattr(obj, 'post', None)


# ---------------------------------------------------------------------------
# post-hooks (skge/param.py:161-174)
# ---------------------------------------------------------------------------


def _rows_post(M, idx, code):
    rowlen = int(np.prod(M.shape[1:]))
    if idx is None:
        _ext.check(_ext.lib().skge_rows_post(_ext.ptr(M.data), None, M.shape[0], rowlen, code, _ext.stream()))
    else:
        it = _as_tensor(idx, torch.int32)
        _ext.check(_ext.lib().skge_rows_post(_ext.ptr(M.data), _ext.ptr(it), it.numel(), rowlen, code,
                                             _ext.stream()))
    return M


def normalize(M, idx=None):
    """Unit-L2 rows (all rows, or rows ``idx`` in place).  skge/param.py:161-167."""
    if isinstance(M, Parameter):
        if idx is None:     # whole table, at creation: plain tensor arithmetic
            M.data.div_(torch.linalg.vector_norm(M.data, dim=1, keepdim=True))
            return M
        return _rows_post(M, idx, _ext.POST_NORMALIZE)
    if idx is None:
        return M / np.sqrt(np.sum(M ** 2, axis=1))[:, np.newaxis]
    nrm = np.sqrt(np.sum(M[idx, :] ** 2, axis=1))[:, np.newaxis]
    M[idx, :] = M[idx, :] / nrm
    return M


def normless1(M, idx=None):
    """Rows divided by max(1, squared norm).  skge/param.py:170-174.

    With ``idx=None`` the reference's ``M[None]`` makes the sum run over rows,
    so each COLUMN is divided by max(1, sum over rows of x^2); kept as is."""
    if isinstance(M, Parameter):
        if idx is None:
            nrm = (M.data ** 2).sum(dim=0, keepdim=True).clamp_(min=1.0)
            M.data.div_(nrm)
            return M
        return _rows_post(M, idx, _ext.POST_NORMLESS1)
    if idx is None:
        nrm = np.sum(M ** 2, axis=0)[np.newaxis, :]
        nrm[nrm < 1] = 1
        return M / nrm
    nrm = np.sum(M[idx] ** 2, axis=1)[:, np.newaxis]
    nrm[nrm < 1] = 1
    M[idx] = M[idx] / nrm
    return M


normalize.code = _ext.POST_NORMALIZE
normless1.code = _ext.POST_NORMLESS1


def post_code(post):
    """SKGE_POST_* of a built-in hook, None for a user callable."""
    if post is None:
        return _ext.POST_NONE
    return getattr(post, 'code', None)


# ---------------------------------------------------------------------------
# updaters (skge/param.py:108-158)
# ---------------------------------------------------------------------------


class ParameterUpdate(object):

    opt_code = None

    def __init__(self, param, learning_rate):
        self.param = param
        self.learning_rate = learning_rate

    def __call__(self, gradient, idx=None):
        code = post_code(self.param.post)
        if code is not None and self.opt_code is not None:
            self._apply(gradient, idx, code)          # update + post-hook in one kernel
            return
        self._update(gradient, idx)
        if self.param.post is not None:
            r = self.param.post(self.param, idx)
            if r is not None:
                self.param = r

    def reset(self):
        pass

    def _state(self):
        return None

    def _apply(self, g, idx, post):
        p = self.param
        gt = _as_tensor(g, torch.float32)
        rowlen = int(np.prod(p.shape[1:]))
        if idx is None:
            it, U = None, p.shape[0]
        else:
            it = _as_tensor(idx, torch.int32)
            U = it.numel()
        if U == 0:
            return
        st = self._state()
        counts = p._update_counts if self.opt_code == _ext.OPT_ADAGRAD else None
        _ext.check(_ext.lib().skge_sparse_update(_ext.ptr(p.data), _ext.ptr(st), _ext.ptr(gt), _ext.ptr(it),
                                                 U, rowlen, self.opt_code, float(self.learning_rate), post,
                                                 _ext.ptr(counts), _ext.stream()))

    def _update(self, g, idx=None):
        self._apply(g, idx, _ext.POST_NONE)


class SGD(ParameterUpdate):
    """param[idx] -= learning_rate * g  (skge/param.py:124-130)."""
    opt_code = _ext.OPT_SGD


class AdaGrad(ParameterUpdate):
    """AdaGrad (skge/param.py:134-158): p2[idx] += g*g; H = max(sqrt(p2[idx]), 1e-7);
    param[idx] -= learning_rate * g / H."""
    opt_code = _ext.OPT_ADAGRAD

    def __init__(self, param, learning_rate):
        super(AdaGrad, self).__init__(param, learning_rate)
        self.p2 = torch.zeros_like(param.data)

    def _state(self):
        return self.p2

    def reset(self):
        self.p2.zero_()     # in place: fused steps captured in a CUDA graph keep this address
