"""Model base class and the stochastic trainers (reference: skge/base.py:1140-1427).

Same constructor keywords, attributes and callback protocol as the reference.
Two execution paths:

* fused (default whenever the model, the updater class and the sampler are the
  built-in ones): the training triples are marshalled ONCE into int32 SoA device
  arrays at ``fit`` time and every minibatch is sample -> score -> mask ->
  segmented mean -> sparse update on the stream, with one host read
  (``nviolations`` / ``loss``) per epoch;
* hook path (custom ``_gradients`` / ``param_update`` / ``samplef``): the
  reference's control flow verbatim, calling the same kernels through the
  model's ``_gradients`` / ``_pairwise_gradients`` and the updaters.
"""
import pickle
import timeit

import logging

import numpy as np
import torch

from . import _ext, kernels
from .param import Parameter, AdaGrad, SGD, post_code

log = logging.getLogger('EX-KG')

_DEF_NBATCHES = 100
_DEF_POST_EPOCH = []
_DEF_LEARNING_RATE = 0.1
_DEF_SAMPLE_FUN = None
_DEF_MAX_EPOCHS = 1000
_DEF_MARGIN = 1.0
# Deviation from the reference (skge/base.py:35-36, 1336-1346): a bare
# PairwiseStochasticTrainer(model) does NOT open gradients.txt / embeddings.txt
# in the working directory; pass file_grad= / file_embed= to get the dumps.
_FILE_GRADIENTS = None
_FILE_EMBEDDINGS = None


class Model(object):
    """Base class of the knowledge-graph models (skge/base.py:1140-1192).

    Subclasses implement ``_scores(ss, ps, os)``, ``_gradients(xys)`` and / or
    ``_pairwise_gradients(pxs, nxs)``.
    """

    track_counters = True   # per-entity instrumentation counters of the fork (cheap device atomics)

    def __init__(self, *args, **kwargs):
        self.params = {}
        self.hyperparams = {}
        self.add_hyperparam('init', kwargs.pop('init', 'nunif'))

    def add_param(self, param_id, shape, post=None, value=None):
        if value is None:
            value = Parameter(shape, self.init, name=param_id, post=post)
        setattr(self, param_id, value)
        self.params[param_id] = value

    def add_hyperparam(self, param_id, value):
        setattr(self, param_id, value)
        self.hyperparams[param_id] = value

    def __getstate__(self):
        return {'hyperparams': self.hyperparams, 'params': self.params}

    _posts = {}             # param id -> post-hook, for pickles written by the reference (which lose it)

    def __setstate__(self, st):
        self.params = {}
        self.hyperparams = {}
        for pid, p in st['params'].items():
            if not isinstance(p, Parameter):      # an ndarray from a reference-format pickle
                p = Parameter.from_reference(p)
                p.name = p.name or pid
                p.post = p.post or self._posts.get(pid)
            self.add_param(pid, None, None, value=p)
        for pid, p in st['hyperparams'].items():
            self.add_hyperparam(pid, p)

    def save(self, fname, protocol=2):
        """Pickle in the REFERENCE's format (skge/base.py:1170-1187): the stream names the
        same classes (skge.hole.HolE, skge.param.Parameter, ...) and stores the parameters as
        float64 ndarray payloads, so the reference -- and its analysis scripts -- can load
        a model trained here, and ``Model.load`` here reads models written by the reference."""
        with open(fname, 'wb') as fout:
            fout.write(dumps_reference(self, protocol))

    @staticmethod
    def load(fname):
        with open(fname, 'rb') as fin:
            return loads_reference(fin.read())


class _RefPickler(pickle.Pickler):
    """Writes our Parameter objects as the reference's ndarray-subclass payload."""

    def reducer_override(self, obj):
        if isinstance(obj, Parameter):
            a = np.ascontiguousarray(np.asarray(obj, dtype=np.float64))
            fn, args, state = a.__reduce__()      # (_reconstruct, (ndarray, (0,), b'b'), state)
            return fn, (Parameter,) + tuple(args[1:]), state
        return NotImplemented


class _RefUnpickler(pickle.Unpickler):

    def find_class(self, module, name):
        if module == 'skge.param' and name == 'Parameter':
            from .param import RefParameter
            return RefParameter
        return super(_RefUnpickler, self).find_class(module, name)


def dumps_reference(obj, protocol=2):
    """Reference-compatible pickle bytes (protocol <= 3, so module names are plain text)."""
    import io
    if protocol > 3:
        protocol = 3
    buf = io.BytesIO()
    _RefPickler(buf, protocol=protocol).dump(obj)
    # numpy >= 2 names its reconstructor numpy._core.multiarray; older numpy (what a
    # reference installation may run) only knows numpy.core.multiarray, which both resolve.
    return buf.getvalue().replace(b'cnumpy._core.multiarray\n', b'cnumpy.core.multiarray\n')


def loads_reference(data):
    import io
    return _RefUnpickler(io.BytesIO(data)).load()


def _check_ids(xs, sz):
    """Out-of-range ids would be an IndexError in the reference (fancy indexing); the kernels
    do not bounds-check, so validate once per fit on the host."""
    if sz is None or len(xs) == 0:
        return
    a = xs if isinstance(xs, np.ndarray) else np.asarray(xs, dtype=np.int64).reshape(-1, 3)
    if a.min() < 0 or a[:, :2].max() >= sz[0] or a[:, 2].max() >= sz[2]:
        raise IndexError('triple ids out of range for sz=%r' % (tuple(sz),))


def _triples_to_device(xs):
    """list / array of (s, o, p) -> three int32 CUDA tensors (SoA)."""
    if isinstance(xs, torch.Tensor):
        a = xs.to(_ext.device())
        return tuple(a[:, i].to(torch.int32).contiguous() for i in range(3))
    a = np.asarray(xs, dtype=np.int64).reshape(-1, 3)
    t = torch.from_numpy(np.ascontiguousarray(a.T, dtype=np.int32)).to(_ext.device())
    return t[0].contiguous(), t[1].contiguous(), t[2].contiguous()


def _opt_code(pu):
    """SKGE_OPT_* when the updater CLASS is exactly the built-in SGD / AdaGrad."""
    if pu is SGD:
        return _ext.OPT_SGD
    if pu is AdaGrad:
        return _ext.OPT_ADAGRAD
    return None


class _GraphedStep(object):
    """Runs ``body(idx)`` -- one fused minibatch -- through a CUDA graph.

    Minibatches of configs 1-4 are a few thousand pairs: ~16 short kernels whose launch gaps
    dominate.  The first call for a given minibatch length runs eagerly (sizes the workspace),
    the second is captured, later ones copy the new example indices into the static index
    buffer and replay.  Everything the step reads besides ``idx`` lives at fixed addresses
    (parameters, optimiser state, counters, the sampler's device-side Philox offset).  Any
    capture failure falls back to eager launches."""

    def __init__(self, body, enabled=True, signature=None):
        self.body, self.enabled, self.slots = body, enabled, {}
        # host scalars and buffer addresses that a captured graph bakes in (learning rate, margin,
        # AdaGrad state pointers ...): when they change the graphs are dropped and captured again
        self.signature = signature
        self._sig = signature() if signature else None

    def __call__(self, batch):
        if self.signature is not None:
            sig = self.signature()
            if sig != self._sig:
                self._sig, self.slots = sig, {}
        B = batch.numel()
        slot = self.slots.get(B)
        if slot is None:
            slot = self.slots[B] = dict(idx=torch.empty(B, dtype=torch.int32, device=batch.device), graph=None,
                                        warm=0, launches=0)
        slot['idx'].copy_(batch)
        if slot['graph'] is not None:
            slot['graph'].replay()
            kernels.LAUNCHES['n'] += slot['launches']
            return
        if not self.enabled or slot['warm'] < 1:
            self.body(slot['idx'])
            slot['warm'] += 1
            return
        try:
            n0 = kernels.LAUNCHES['n']
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.body(slot['idx'])
            slot['launches'] = kernels.LAUNCHES['n'] - n0
            slot['graph'] = g
            # the captured kernels hold raw pointers into the shared scratch buffer: keep that very
            # tensor alive even if a later, larger call makes the library wrapper allocate a new one
            slot['keep'] = kernels._ws.buf
            g.replay()          # capturing records the work, it does not run it
        except Exception as e:  # noqa: BLE001 -- e.g. a driver that refuses the capture
            log.warning('CUDA graph capture failed (%s); continuing with eager launches', e)
            self.enabled = False
            torch.cuda.synchronize()
            self.body(slot['idx'])


class StochasticTrainer(object):
    """Stochastic gradient descent trainer with scalar (logistic) loss
    (skge/base.py:1195-1316).  Models implement ``_gradients(xys)``."""

    def __init__(self, *args, **kwargs):
        self.model = args[0]
        self.hyperparams = {}
        self.add_hyperparam('max_epochs', kwargs.pop('max_epochs', _DEF_MAX_EPOCHS))
        self.add_hyperparam('nbatches', kwargs.pop('nbatches', _DEF_NBATCHES))
        self.add_hyperparam('learning_rate', kwargs.pop('learning_rate', _DEF_LEARNING_RATE))
        self.post_epoch = kwargs.pop('post_epoch', _DEF_POST_EPOCH)
        self.samplef = kwargs.pop('samplef', _DEF_SAMPLE_FUN)
        pu = kwargs.pop('param_update', AdaGrad)
        self._param_update = pu
        self._updaters = {key: pu(param, self.learning_rate) for key, param in self.model.params.items()}
        self.seed = kwargs.pop('seed', 42)       # the reference seeds numpy with 42 at import
        self.fused = kwargs.pop('fused', True)   # set False to force the reference's hook path
        self.cuda_graphs = kwargs.pop('cuda_graphs', True)   # replay fused minibatches as CUDA graphs
        self._gen = None

    def set_max_epochs(self, epoch):
        self.max_epochs = epoch

    def __getstate__(self):
        return self.hyperparams

    def __setstate__(self, st):
        self.hyperparams = {}
        for pid, p in st.items():
            self.add_hyperparam(pid, p)

    def add_hyperparam(self, param_id, value):
        setattr(self, param_id, value)
        self.hyperparams[param_id] = value

    # -- shared machinery ---------------------------------------------------------
    def _device_sampler(self):
        """The built-in sampler object behind ``samplef`` (None if user code)."""
        from .sample import Sampler
        owner = getattr(self.samplef, '__self__', None)
        if isinstance(owner, Sampler) and getattr(self.samplef, '__func__', None) is Sampler.sample \
                and owner.device_ready():
            return owner
        return None

    def _can_fuse(self, fused_attr):
        if not self.fused or _opt_code(self._param_update) is None:
            return False
        if not hasattr(self.model, fused_attr) or hasattr(self.model, '_prepare_batch_step'):
            return False
        # a subclass that overrides one of the reference's extension points must see it called
        owner = next((c for c in type(self.model).__mro__ if fused_attr in c.__dict__), None)
        for hook in ('_scores', '_gradients', '_pairwise_gradients'):
            if getattr(type(self.model), hook, None) is not getattr(owner, hook, None):
                return False
        for base in (StochasticTrainer, PairwiseStochasticTrainer):
            if isinstance(self, base):
                own = base
        if type(self)._process_batch is not own._process_batch or type(self)._batch_step is not own._batch_step:
            return False
        if any(post_code(p.post) is None for p in self.model.params.values()):
            return False
        return self.samplef is None or self._device_sampler() is not None

    def _graph_signature(self):
        """Everything a captured minibatch graph freezes besides the example indices."""
        m = self.model
        sig = [self.learning_rate, getattr(m, 'margin', None), getattr(m, 'rparam', None),
               getattr(getattr(m, 'af', None), '__name__', None), getattr(m, 'l1', None)]
        for key, u in self._updaters.items():
            st = u._state() if hasattr(u, '_state') else None
            sig += [key, u.learning_rate, u.param.data.data_ptr(), st.data_ptr() if st is not None else 0]
        return tuple(sig)

    def _randperm(self, n):
        if self._gen is None:
            self._gen = torch.Generator(device=_ext.device())
            self._gen.manual_seed(int(self.seed))
        return torch.randperm(n, device=_ext.device(), generator=self._gen, dtype=torch.int64)

    def _batch_bounds(self, n):
        """nbatches slices of n // nbatches plus a remainder slice when
        n % nbatches != 0 (skge/base.py:1246-1252, 1268)."""
        self.batch_size = n // self.nbatches
        cuts = list(range(self.batch_size, n, self.batch_size))
        return list(zip([0] + cuts, cuts + [n]))

    def _run_epochs(self, n, step):
        """The epoch loop of ``_optim`` (skge/base.py:1254-1291); ``step(batch)``
        gets an int32 CUDA tensor of example indices."""
        bounds = self._batch_bounds(n)
        prepare = getattr(self.model, '_prepare_fused', None)
        for self.epoch in range(1, self.max_epochs + 1):
            self._pre_epoch()
            self.epoch_start = timeit.default_timer()
            if prepare is not None:
                prepare()       # per-epoch derived state of the model (HolE: spectra of E, R): part of the epoch
            perm = self._randperm(n).to(torch.int32)
            for lo, hi in bounds:
                step(perm[lo:hi])
            self._end_epoch()
            for f in self.post_epoch:
                if not f(self):
                    break   # as in the reference, this only leaves the callback loop
        if hasattr(self.model, '_end_fused'):
            self.model._end_fused()

    # -- reference API ----------------------------------------------------------------
    def fit(self, xs, ys):
        if self._can_fuse('_fused_logistic_step'):
            self._fit_fused(xs, ys)
        else:
            self._optim(list(zip(xs, ys)))

    def _pre_epoch(self):
        self.loss = 0
        if getattr(self, '_loss_dev', None) is not None:
            self._loss_dev.zero_()

    def _end_epoch(self):
        if getattr(self, '_loss_dev', None) is not None:
            self.loss = float(self._loss_dev.item())

    def _fit_fused(self, xs, ys):
        dev = _ext.device()
        _check_ids(xs, getattr(self.model, 'sz', None))
        s, o, p = _triples_to_device(xs)
        y = _ext.as_f32(np.asarray(ys, dtype=np.float32))
        n = s.numel()
        self._loss_dev = torch.zeros(1, dtype=torch.float64, device=dev)
        self._counts = torch.zeros(4, dtype=torch.int32, device=dev)
        sampler = self._device_sampler() if self.samplef is not None else None
        philox = torch.zeros(1, dtype=torch.int64, device=dev)
        per_pos = sampler.n * len(sampler.modes) if sampler is not None else 0

        bufs = {}   # per minibatch length: positives followed by their negatives (captured graphs keep the pointers)
        # (s, o, p, y) as the rows of one table: a minibatch's positives are ONE column gather
        soa = torch.stack([s, o, p, y.view(torch.int32)]) if sampler is not None else None

        def body(idx):
            bl = idx.long()
            nb = idx.numel()
            if sampler is None:
                self.model._fused_logistic_step(self._updaters, s[bl].contiguous(), o[bl].contiguous(),
                                                p[bl].contiguous(), y[bl].contiguous(), self._counts,
                                                self._loss_dev, valid=None)
                return
            # xys += samplef(xys): positives, then their negatives labelled -1 (skge/base.py:1295-1296);
            # negatives that exhausted their tries are masked out instead of compacted (no host sync).
            # The sampler writes the negatives straight into the tail of the minibatch arrays.
            if nb not in bufs:
                nneg = nb * per_pos
                tab = torch.empty(4, nb + nneg, dtype=torch.int32, device=dev)     # rows s, o, p, y (bit pattern)
                tab[3].view(torch.float32).fill_(-1.0)
                bufs[nb] = (tab, torch.ones(nb + nneg, dtype=torch.uint8, device=dev),
                            [torch.empty(nneg, dtype=torch.int32, device=dev) for _ in range(3)])
            tab, valid, scratch = bufs[nb]
            bs, bo, bp, by = tab[0], tab[1], tab[2], tab[3].view(torch.float32)
            tab[:, :nb].copy_(soa.index_select(1, bl))
            sampler.device_sample(None, nb, 0, src=(bs[:nb], bo[:nb], bp[:nb]), offset_dev=philox,
                                  outs=scratch + [bs[nb:], bo[nb:], bp[nb:]], valid=valid[nb:])
            philox.add_(nb * per_pos)
            self.model._fused_logistic_step(self._updaters, bs, bo, bp, by, self._counts, self._loss_dev, valid=valid)

        self._run_epochs(n, _GraphedStep(body, self.cuda_graphs, self._graph_signature))

    def _optim(self, xys):
        """Hook path: the reference's loop on host lists (skge/base.py:1242-1291)."""
        idx = np.arange(len(xys))
        self.batch_size = len(xys) // self.nbatches
        batch_idx = np.arange(self.batch_size, len(xys), self.batch_size)
        rng = np.random.RandomState(self.seed)
        for self.epoch in range(1, self.max_epochs + 1):
            self._pre_epoch()
            rng.shuffle(idx)
            self.epoch_start = timeit.default_timer()
            for batch in np.split(idx, batch_idx):
                bxys = [xys[z] for z in batch]
                self._process_batch(bxys)
            for f in self.post_epoch:
                if not f(self):
                    break

    def _process_batch(self, xys):
        if self.samplef is not None:
            xys += self.samplef(xys)
        if hasattr(self.model, '_prepare_batch_step'):
            self.model._prepare_batch_step(xys)
        grads = self.model._gradients(xys)
        self.loss += self.model.loss
        self._batch_step(grads)

    def _batch_step(self, grads):
        for paramID in self._updaters.keys():
            self._updaters[paramID](*grads[paramID])


class PairwiseStochasticTrainer(StochasticTrainer):
    """Stochastic gradient descent trainer with pairwise ranking loss
    (skge/base.py:1320-1427).  Models implement ``_pairwise_gradients(pxs, nxs)``."""

    def __init__(self, *args, **kwargs):
        margin = kwargs.pop('margin', _DEF_MARGIN)
        fg = kwargs.pop('file_grad', _FILE_GRADIENTS)
        fe = kwargs.pop('file_embed', _FILE_EMBEDDINGS)
        super(PairwiseStochasticTrainer, self).__init__(*args, **kwargs)
        self.model.add_hyperparam('margin', margin)   # stored on the MODEL (skge/base.py:1335)
        self.file_gradients = open(fg, 'w') if fg is not None else None
        self.file_embeddings = open(fe, 'w') if fe is not None else None
        self.pickle_file_embeddings = open(fe + '.pkl', 'wb') if fe is not None else None
        self._nviol_dev = None

    def fit(self, xs, ys):
        fused = self._can_fuse('_fused_pair_step')
        if self.samplef is None:
            # supplied-negatives mode (skge/base.py:1350-1357)
            ysa = np.asarray(ys)
            pidx = np.where(ysa == 1)[0]
            nidx = np.where(ysa != 1)[0]
            pxs = [xs[i] for i in pidx]
            self.nxs = [xs[i] for i in nidx]
            self.pxs = int(len(self.nxs) / len(pxs)) * pxs
            n = min(len(pxs), len(self.nxs))
            if fused:
                self._fit_fused_supplied(n)
            else:
                self._optim(list(range(n)))
            return
        if fused:
            self._fit_fused_sampled(xs)
        else:
            self._optim(list(zip(xs, ys)))
        self._post_fit(xs)

    # -- fused paths ---------------------------------------------------------------
    def _setup_fused(self):
        dev = _ext.device()
        self._nviol_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._counts = torch.zeros(4, dtype=torch.int32, device=dev)

    def _fit_fused_sampled(self, xs):
        _check_ids(xs, getattr(self.model, 'sz', None))
        self._setup_fused()
        sampler = self._device_sampler()
        sampler.ensure_device()
        # the positives are fit's xs; the sampler's own triple set only rejects negatives
        # (skge/base.py:1394-1402, skge/sample.py:28-46) -- the two may differ, e.g. a sampler built
        # on train + valid + test
        src = _triples_to_device(xs)
        n = src[0].numel()
        philox = torch.zeros(1, dtype=torch.int64, device=_ext.device())   # device-side draw counter
        per_pos = sampler.n * len(sampler.modes)

        def body(idx):
            pos, neg, valid = sampler.device_sample(idx, idx.numel(), 0, src=src, offset_dev=philox)
            philox.add_(idx.numel() * per_pos)
            self.model._fused_pair_step(self._updaters, pos, neg, valid, self._counts, self._nviol_dev)

        self._run_epochs(n, _GraphedStep(body, self.cuda_graphs, self._graph_signature))

    def _fit_fused_supplied(self, n):
        _check_ids(self.pxs + self.nxs, getattr(self.model, 'sz', None))
        self._setup_fused()
        P = torch.stack(_triples_to_device(self.pxs), 1)
        Nn = torch.stack(_triples_to_device(self.nxs), 1)
        self._sup = [P, Nn]       # shuffled IN PLACE every epoch (graph-captured steps keep the pointers)

        def body(idx):
            bl = idx.long()
            pp, nn = self._sup[0][bl], self._sup[1][bl]
            pos = tuple(pp[:, i].contiguous() for i in range(3))
            neg = tuple(nn[:, i].contiguous() for i in range(3))
            self.model._fused_pair_step(self._updaters, pos, neg, None, self._counts, self._nviol_dev)

        self._run_epochs(n, _GraphedStep(body, self.cuda_graphs, self._graph_signature))

    def _pre_epoch(self):
        self.nviolations = 0
        if self._nviol_dev is not None:
            self._nviol_dev.zero_()
        if self.samplef is None:
            if getattr(self, '_sup', None) is not None:
                # independent shuffles of positives and negatives (skge/base.py:1390-1392)
                for t in self._sup:
                    t.copy_(t[self._randperm(t.shape[0])])
            else:
                rng = np.random.RandomState(self.seed + getattr(self, 'epoch', 0))
                rng.shuffle(self.pxs)
                rng.shuffle(self.nxs)

    def _end_epoch(self):
        if self._nviol_dev is not None:
            self.nviolations = int(self._nviol_dev.item())

    # -- hook path ---------------------------------------------------------------------
    def _process_batch(self, xys):
        pxs, nxs = [], []
        for xy in xys:
            if self.samplef is not None:
                for nx in self.samplef([xy]):
                    pxs.append(xy)
                    nxs.append(nx)
            else:
                pxs.append((self.pxs[xy], 1))
                nxs.append((self.nxs[xy], 1))
        if hasattr(self.model, '_prepare_batch_step'):
            self.model._prepare_batch_step(pxs, nxs)
        grads = self.model._pairwise_gradients(pxs, nxs)
        if grads is not None:
            self.nviolations += self.model.nviolations
            self._batch_step(grads)

    # -- end-of-fit instrumentation (skge/base.py:1364-1386) ------------------------------
    def _post_fit(self, xs):
        E = self.model.params.get('E')
        if E is None:
            return
        s, o, _ = _triples_to_device(xs)
        n = E.shape[0]
        E._neighbours += (torch.bincount(s.long(), minlength=n) + torch.bincount(o.long(), minlength=n)).int()
        if self.file_gradients is not None:
            self.file_gradients.write('Entity,Degree,#(violations),#(updates)\n')
            for i, (en, ev, ec) in enumerate(zip(E.neighbours, E.violations, E.updateCounts)):
                self.file_gradients.write('%d,%d,%d,%d\n' % (i, en, ev, ec))
            self.file_gradients.flush()
        if self.file_embeddings is not None or self.pickle_file_embeddings is not None:
            emb = np.asarray(E, dtype=np.float64)
            if self.file_embeddings is not None:
                for i, e in enumerate(emb):
                    self.file_embeddings.write('%d,%s\n' % (i, str(e)))
                self.file_embeddings.flush()
            if self.pickle_file_embeddings is not None:
                pickle.dump(list(emb), self.pickle_file_embeddings, protocol=2)
                self.pickle_file_embeddings.flush()


# the evaluator and metric helpers live in skge/base.py in the reference
from .ranking import (FilteredRankingEval, LinkPredictionEval, ranking_scores, compute_scores,  # noqa: E402,F401
                      _print_pos)
