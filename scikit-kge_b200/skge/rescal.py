"""RESCAL on the device (reference: skge/rescal.py)."""

from . import _ext, kernels
from . import actfun as af
from .base import Model
from .param import DevArray, post_code
from ._modelutil import idx_tensor, unzip_device, updater_args


class RESCAL(Model):
    """RESCAL(sz, ncomp, rparam=0.0, af='linear', init='nunif') -- skge/rescal.py:19-29.

    score(s, p, o) = E[s]^T W[p] E[o] (skge/rescal.py:31-35).  Training with
    the logistic loss (StochasticTrainer) is supported.  The reference's
    ``_pairwise_gradients`` raises on numpy >= 1.24 (ragged array at
    skge/rescal.py:84) and is not part of any benchmark configuration; it is not
    provided here either.
    """

    model_code = _ext.MODEL_RESCAL

    def __init__(self, *args, **kwargs):
        super(RESCAL, self).__init__(*args, **kwargs)
        self.add_hyperparam('sz', args[0])
        self.add_hyperparam('ncomp', args[1])
        self.add_hyperparam('rparam', kwargs.pop('rparam', 0.0))
        aff = kwargs.pop('af', 'linear')
        self.add_hyperparam('af', af.afuns[aff])
        self.add_param('E', (self.sz[0], self.ncomp))
        self.add_param('W', (self.sz[2], self.ncomp, self.ncomp))
        self.track_counters = kwargs.pop('track_counters', True)

    def _scores(self, ss, ps, os):
        out = kernels.scores(self.model_code, self.E.data, self.W.data, idx_tensor(ss), idx_tensor(ps),
                             idx_tensor(os))
        return out.cpu().numpy()

    def _gradients(self, xys):
        """{'E': (ge, eidx), 'W': (gw, pidx)}; sets ``loss`` (skge/rescal.py:37-76)."""
        s, o, p, y = unzip_device(xys, with_ys=True)
        r = kernels.logistic_grads(self.model_code, self.E.data, self.W.data, s, o, p, y, self.rparam)
        self.loss = r['loss']
        return {'E': (DevArray(r['ge']), DevArray(r['eidx'])), 'W': (DevArray(r['g2']), DevArray(r['idx2']))}

    def _pairwise_gradients(self, pxs, nxs):
        raise NotImplementedError('RESCAL pairwise training is broken in the reference '
                                  '(skge/rescal.py:84) and not provided')

    def _fused_logistic_step(self, updaters, s, o, p, y, counts, loss_accum, valid=None):
        opt, lr, p2E, p2W = updater_args(updaters, 'E', 'W')
        tc = self.track_counters and opt == _ext.OPT_ADAGRAD
        kernels.logistic_step(self.model_code, self.E.data, self.W.data, p2E, p2W, s, o, p, y, self.rparam, opt,
                              lr, post_code(self.E.post), post_code(self.W.post), counts, loss_accum,
                              self.E._update_counts if tc else None, self.W._update_counts if tc else None, valid=valid)
