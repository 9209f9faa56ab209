"""TransE on the device (reference: skge/transe.py)."""
import logging


from . import _ext, kernels
from .base import Model
from .param import normalize, DevArray, post_code
from ._modelutil import idx_tensor, unzip_device, updater_args

log = logging.getLogger('EX-KG')


class TransE(Model):
    """Translational Embeddings of Knowledge Graphs.

    TransE(sz, ncomp, l1=True, init='nunif') -- skge/transe.py:14-23.
    score(s, p, o) = -||E[s] + R[p] - E[o]||_1, or minus the SQUARED L2
    distance when l1=False (skge/transe.py:25-46).
    """

    _posts = {'E': normalize}
    model_code = _ext.MODEL_TRANSE

    def __init__(self, *args, **kwargs):
        super(TransE, self).__init__(*args, **kwargs)
        self.add_hyperparam('sz', args[0])
        self.add_hyperparam('ncomp', args[1])
        self.add_hyperparam('l1', kwargs.pop('l1', True))
        self.add_param('E', (self.sz[0], self.ncomp), post=normalize)
        self.add_param('R', (self.sz[2], self.ncomp))
        self.track_counters = kwargs.pop('track_counters', True)
        log.info("l1 is %r " % (self.l1))

    def _scores(self, ss, ps, os):
        out = kernels.scores(self.model_code, self.E.data, self.R.data, idx_tensor(ss), idx_tensor(ps),
                             idx_tensor(os), l1=self.l1)
        return out.cpu().numpy()

    def _pairwise_gradients(self, pxs, nxs):
        """{'E': (ge, eidx), 'R': (gr, ridx)} or None when no pair violates the
        margin; sets ``nviolations`` (skge/transe.py:48-165)."""
        pos, neg = unzip_device(pxs), unzip_device(nxs)
        r = kernels.pair_grads(self.model_code, self.E.data, self.R.data, pos, neg, None, self.margin,
                               self.l1, ent_viol=self.E._violations if self.track_counters else None)
        self.nviolations = r['nviol']
        self.last_scores = (r['pscores'], r['nscores'])
        if r['nviol'] == 0:
            return
        return {'E': (DevArray(r['ge']), DevArray(r['eidx'])), 'R': (DevArray(r['gr']), DevArray(r['ridx']))}

    def _fused_pair_step(self, updaters, pos, neg, valid, counts, nviol_accum):
        opt, lr, p2E, p2R = updater_args(updaters, 'E', 'R')
        tc = self.track_counters
        kernels.pair_step(self.model_code, self.E.data, self.R.data, p2E, p2R, pos, neg, valid, self.margin,
                          self.l1, 0.0, opt, lr, post_code(self.E.post), post_code(self.R.post), counts,
                          nviol_accum, ent_viol=self.E._violations if tc else None,
                          ucE=self.E._update_counts if tc and opt == _ext.OPT_ADAGRAD else None,
                          ucR=self.R._update_counts if tc and opt == _ext.OPT_ADAGRAD else None)
