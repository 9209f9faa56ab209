"""Activation functions of the pairwise / logistic losses.

Same public surface as the reference's skge/actfun.py:13-73 -- classes ``Linear``,
``Sigmoid``, ``Tanh``, ``ReLU``, ``Softplus`` with static ``f`` / ``g_given_f`` (derivative
expressed through the function value), ``key()`` and the ``afuns`` registry keyed by the
lower-cased class name -- but table-driven: every entry also carries the ``code`` the CUDA
kernels switch on (SKGE_AF_* in include/skge_b200.h; device versions in csrc/common.cuh).
The numpy callables here serve host-side callers only; training never goes through them.
"""
import numpy as np

from . import _ext


class ActivationFunction(object):
    """Base of the generated classes (isinstance / subclass checks keep working)."""
    code = None

    @classmethod
    def key(cls):
        return cls.__name__.lower()


def _no_derivative(_fx):
    raise NotImplementedError()


#  name        kernel code        f(x)                                  g_given_f(f(x))
_SPEC = (
    ('Linear',   _ext.AF_LINEAR,  lambda x: x,                          lambda fx: np.ones(fx.shape[0])),
    ('Sigmoid',  _ext.AF_SIGMOID, lambda x: 1.0 / (1 + np.exp(-x)),     lambda fx: fx * (1.0 - fx)),
    ('Tanh',     _ext.AF_TANH,    np.tanh,                              lambda fx: 1 - fx ** 2),
    ('ReLU',     _ext.AF_RELU,    lambda x: np.maximum(0, x),           lambda fx: np.int_(fx > 0)),
    # the reference defines no g_given_f for Softplus either (it only has a raising ``g``)
    ('Softplus', None,            lambda x: np.log(1 + np.exp(x)),      None),
)

afuns = {}
for _name, _code, _f, _g in _SPEC:
    _attrs = {'code': _code, 'f': staticmethod(_f), '__doc__': '%s activation (kernel code %r)' % (_name, _code)}
    if _g is not None:
        _attrs['g_given_f'] = staticmethod(_g)
    else:
        _attrs['g'] = staticmethod(_no_derivative)
    _cls = type(_name, (ActivationFunction,), _attrs)
    _cls.__module__ = __name__
    globals()[_name] = _cls
    afuns[_cls.key()] = _cls
del _name, _code, _f, _g, _attrs, _cls


def af_code(af):
    """Kernel enum of an activation given as class or registry key."""
    if isinstance(af, str):
        af = afuns[af]
    code = getattr(af, 'code', None)
    if code is None:
        raise NotImplementedError('activation %r has no g_given_f (as in the reference)' % (af,))
    return code
