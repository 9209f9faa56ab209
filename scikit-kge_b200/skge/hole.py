"""HolE on the device (reference: skge/hole.py)."""
import torch

from . import _ext, kernels
from . import actfun as af
from .base import Model
from .param import normless1, DevArray, post_code
from ._modelutil import idx_tensor, unzip_device, updater_args


def spectral_len_ok(d):
    """Row lengths the device transforms take (csrc/fft.cuh): even, 32..1024, d / 2 = 2^a 3^b 5^c
    -- every power of two, and e.g. 100, 150, 200, 300."""
    if d < 32 or d > 1024 or d % 2:
        return False
    h = d // 2
    for r in (2, 3, 5):
        while h % r == 0:
            h //= r
    return h == 1


class HolE(Model):
    """Holographic embeddings.

    HolE(sz, ncomp, rparam=0.0, af=Sigmoid, init='nunif') -- skge/hole.py:10-17.
    score(s, p, o) = sum_k R[p]_k ccorr(E[s], E[o])_k (skge/hole.py:19-20); the
    circular correlations run in shared memory instead of numpy's FFT.
    """

    _posts = {'E': normless1}
    model_code = _ext.MODEL_HOLE

    def __init__(self, *args, **kwargs):
        super(HolE, self).__init__(*args, **kwargs)
        self.add_hyperparam('sz', args[0])
        self.add_hyperparam('ncomp', args[1])
        self.add_hyperparam('rparam', kwargs.pop('rparam', 0.0))
        self.add_hyperparam('af', kwargs.pop('af', af.Sigmoid))
        self.add_param('E', (self.sz[0], self.ncomp), post=normless1)
        self.add_param('R', (self.sz[2], self.ncomp))
        self.track_counters = kwargs.pop('track_counters', True)

    def _scores(self, ss, ps, os):
        out = kernels.scores(self.model_code, self.E.data, self.R.data, idx_tensor(ss), idx_tensor(ps),
                             idx_tensor(os))
        return out.cpu().numpy()

    def _gradients(self, xys):
        """Logistic-loss gradients; sets ``loss`` (skge/hole.py:22-42)."""
        s, o, p, y = unzip_device(xys, with_ys=True)
        r = kernels.logistic_grads(self.model_code, self.E.data, self.R.data, s, o, p, y, self.rparam)
        self.loss = r['loss']
        return {'E': (DevArray(r['ge']), DevArray(r['eidx'])), 'R': (DevArray(r['g2']), DevArray(r['idx2']))}

    def _pairwise_gradients(self, pxs, nxs):
        """Pairwise-margin gradients, None when nothing violates; sets
        ``nviolations`` (skge/hole.py:44-100)."""
        pos, neg = unzip_device(pxs), unzip_device(nxs)
        r = kernels.pair_grads(self.model_code, self.E.data, self.R.data, pos, neg, None, self.margin,
                               af.af_code(self.af), self.rparam)
        self.nviolations = r['nviol']
        self.last_scores = (r['pscores'], r['nscores'])
        if r['nviol'] == 0:
            return
        return {'E': (DevArray(r['ge']), DevArray(r['eidx'])), 'R': (DevArray(r['gr']), DevArray(r['ridx']))}

    def _uc(self, opt):
        tc = self.track_counters and opt == _ext.OPT_ADAGRAD
        return (self.E._update_counts if tc else None), (self.R._update_counts if tc else None)

    # -- frequency-domain training state (even ncomp with ncomp / 2 = 2^a 3^b 5^c) -----------
    spectral = True         # set False to force the per-pair FFT / direct kernels

    def _prepare_fused(self):
        """Called by the trainers at the start of every fused epoch: (re)build the packed
        spectra of E and R that the spectral step reads and keeps current."""
        d = self.ncomp
        if self.spectral and spectral_len_ok(d):
            old = getattr(self, '_spec', None)      # refresh in place: captured graphs keep the pointers
            if old is None or old[0].shape != self.E.data.shape or old[0].device != self.E.data.device:
                old = (torch.empty_like(self.E.data), torch.empty_like(self.R.data))
            self._spec = (kernels.hole_spectra(self.E.data, out=old[0]), kernels.hole_spectra(self.R.data, out=old[1]))
        else:
            self._spec = None

    def _end_fused(self):
        self._spec_keep, self._spec = getattr(self, '_spec', None), None

    def _fused_pair_step(self, updaters, pos, neg, valid, counts, nviol_accum):
        opt, lr, p2E, p2R = updater_args(updaters, 'E', 'R')
        ucE, ucR = self._uc(opt)
        spec = getattr(self, '_spec', None)
        if spec is not None:
            kernels.hole_pair_step_spectral(self.E.data, self.R.data, spec[0], spec[1], p2E, p2R, pos, neg, valid,
                                            self.margin, af.af_code(self.af), self.rparam, opt, lr,
                                            post_code(self.E.post), post_code(self.R.post), counts, nviol_accum,
                                            ucE=ucE, ucR=ucR)
            return
        kernels.pair_step(self.model_code, self.E.data, self.R.data, p2E, p2R, pos, neg, valid, self.margin,
                          af.af_code(self.af), self.rparam, opt, lr, post_code(self.E.post),
                          post_code(self.R.post), counts, nviol_accum, ucE=ucE, ucR=ucR)

    def _fused_logistic_step(self, updaters, s, o, p, y, counts, loss_accum, valid=None):
        opt, lr, p2E, p2R = updater_args(updaters, 'E', 'R')
        ucE, ucR = self._uc(opt)
        kernels.logistic_step(self.model_code, self.E.data, self.R.data, p2E, p2R, s, o, p, y, self.rparam, opt,
                              lr, post_code(self.E.post), post_code(self.R.post), counts, loss_accum, ucE, ucR, valid=valid)
