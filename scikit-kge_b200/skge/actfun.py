"""Activation functions (reference: skge/actfun.py:13-73).

Host-side classes keep the reference's interface (``f``, ``g_given_f``,
``key()``); ``code`` is the enum the kernels take (SKGE_AF_*).  Softplus has no
``g_given_f`` in the reference either and therefore no kernel code.
"""
import numpy as np

from . import _ext


class ActivationFunction(object):
    code = None

    @classmethod
    def key(cls):
        return cls.__name__.lower()


class Linear(ActivationFunction):
    code = _ext.AF_LINEAR

    @staticmethod
    def f(x):
        return x

    @staticmethod
    def g_given_f(fx):
        return np.ones(fx.shape[0])


class Sigmoid(ActivationFunction):
    code = _ext.AF_SIGMOID

    @staticmethod
    def f(x):
        return 1.0 / (1 + np.exp(-x))

    @staticmethod
    def g_given_f(fx):
        return fx * (1.0 - fx)


class Tanh(ActivationFunction):
    code = _ext.AF_TANH

    @staticmethod
    def f(x):
        return np.tanh(x)

    @staticmethod
    def g_given_f(fx):
        return 1 - fx ** 2


class ReLU(ActivationFunction):
    code = _ext.AF_RELU

    @staticmethod
    def f(x):
        return np.maximum(0, x)

    @staticmethod
    def g_given_f(fx):
        return np.int_(fx > 0)


class Softplus(ActivationFunction):

    @staticmethod
    def f(x):
        return np.log(1 + np.exp(x))

    @staticmethod
    def g(x):
        raise NotImplementedError()


afuns = {}
for cls in ActivationFunction.__subclasses__():
    afuns[cls.key()] = cls


def af_code(af):
    """Kernel enum of an activation given as class or registry key."""
    if isinstance(af, str):
        af = afuns[af]
    code = getattr(af, 'code', None)
    if code is None:
        raise NotImplementedError('activation %r has no g_given_f (as in the reference)' % (af,))
    return code
