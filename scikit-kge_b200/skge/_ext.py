"""ctypes binding of libskge_b200.so (the C ABI declared in include/skge_b200.h).

PyTorch is used for device memory and streams only: every call passes raw
``tensor.data_ptr()`` pointers plus the current CUDA stream.  There is no CPU
fallback: if the library is missing or no sm_100 device is present, the first
use raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libskge_b200.so')

OPT_SGD, OPT_ADAGRAD = 0, 1
POST_NONE, POST_NORMALIZE, POST_NORMLESS1 = 0, 1, 2
AF_LINEAR, AF_SIGMOID, AF_TANH, AF_RELU = 0, 1, 2, 3
MODEL_TRANSE, MODEL_HOLE, MODEL_RESCAL = 0, 1, 2
RANK_L1, RANK_DOT = 0, 1

_P, _I, _L, _F, _Z, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); mirrors include/skge_b200.h one to one
SIGNATURES = {
    'skge_version': (_I, []),
    'skge_last_error': (C.c_char_p, []),
    'skge_check_device': (_I, []),
    'skge_scores_transe': (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _P, _P]),
    'skge_scores_hole': (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P]),
    'skge_scores_rescal': (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P]),
    'skge_pair_workspace_bytes': (_Z, [_L, _I, _I, _L, _L]),
    'skge_transe_pair_grads': (_I, [_P, _P] + [_P] * 7 + [_L, _L, _L, _I, _I, _F] + [_P] * 8 + [_P, _Z, _P]),
    'skge_transe_pair_step': (_I, [_P] * 4 + [_P] * 7 + [_L, _L, _L, _I, _I, _F, _I, _F, _I, _I] + [_P] * 5
                              + [_P, _Z, _P]),
    'skge_hole_pair_grads': (_I, [_P, _P] + [_P] * 7 + [_L, _L, _L, _I, _I, _F, _F] + [_P] * 7 + [_P, _Z, _P]),
    'skge_hole_pair_step': (_I, [_P] * 4 + [_P] * 7 + [_L, _L, _L, _I, _I, _F, _F, _I, _F, _I, _I] + [_P] * 4
                            + [_P, _Z, _P]),
    'skge_hole_spectra': (_I, [_P, _L, _I, _P, _P]),
    'skge_hole_pair_step_spectral': (_I, [_P] * 6 + [_P] * 7 + [_L, _L, _L, _I, _I, _F, _F, _I, _F, _I, _I] + [_P] * 4
                                     + [_P, _Z, _P]),
    'skge_logistic_workspace_bytes': (_Z, [_I, _L, _I, _L, _L]),
    'skge_hole_logistic_grads': (_I, [_P] * 7 + [_L, _L, _L, _I, _F] + [_P] * 6 + [_P, _Z, _P]),
    'skge_hole_logistic_step': (_I, [_P] * 9 + [_L, _L, _L, _I, _F, _I, _F, _I, _I] + [_P] * 4 + [_P, _Z, _P]),
    'skge_rescal_logistic_grads': (_I, [_P] * 7 + [_L, _L, _L, _I, _F] + [_P] * 6 + [_P, _Z, _P]),
    'skge_rescal_logistic_step': (_I, [_P] * 9 + [_L, _L, _L, _I, _F, _I, _F, _I, _I] + [_P] * 4 + [_P, _Z, _P]),
    'skge_sparse_update': (_I, [_P, _P, _P, _P, _L, _L, _I, _F, _I, _P, _P]),
    'skge_rows_post': (_I, [_P, _P, _L, _L, _I, _P]),
    'skge_tripleset_bytes': (_Z, [_L]),
    'skge_tripleset_build': (_I, [_P, _Z, _P, _P, _P, _L, _I, _P]),
    'skge_tripleset_contains': (_I, [_P, _Z, _P, _P, _P, _L, _P, _P]),
    'skge_sample_corrupt': (_I, [_P, _Z, _P, _Z, _P, _P, _P, _P, _L, _I, _I, _L, _L, _I, _U64, _U64, _P]
                            + [_P] * 7 + [_P]),
    'skge_rank_make_queries': (_I, [_I, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F] + [_P] * 5 + [_P]),
    'skge_rank_sweep_packed_floats': (_L, [_L, _I]),
    'skge_rank_sweep_pack': (_I, [_P, _L, _I, _P, _P]),
    'skge_rank_sweep_tiles': (_I, [_I, _P, _L, _L, _I, _P, _P, _P, _L, _P, _P, _P, _L, _P, _P]),
    'skge_rank_rescore': (_I, [_I, _P, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    'skge_rank_scores_one': (_I, [_I, _P, _L, _I, _P, _P, _P]),
    'skge_rank_packed_bytes': (_Z, [_L, _I]),
    'skge_rank_pack_f16': (_I, [_P, _L, _I, _P, _F, _P, _P, _P, _P, _P]),
    'skge_rank_query_scale': (_I, [_P, _P, _P, _L, _I, _F, _P, _P, _P, _P]),
    'skge_rank_gemm_count': (_I, [_P, _P, _L, _L, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P]),
    'skge_rank_quant_lo': (_I, [_P, _L, _I, _P, _P, _P]),
    'skge_rank_quant_lo_s8': (_I, [_P, _L, _I, _P, _P, _P]),
    'skge_rank_pack_q8': (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P, _P]),
    'skge_rank_quant_rows': (_I, [_P, _L, _I, _F, _P, _P, _P, _P]),
    'skge_rank_pack_q8x2': (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P, _P]),
    'skge_rank_single_count': (_I, [_P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _L, _P, _P]),
    'skge_rank_refine_count': (_I, [_P, _P, _P, _P, _P, _L, _L, _P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _L, _P, _P]),
}

_lib = None
_device_ok = False


def load_library():
    """dlopen the C ABI and attach prototypes (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libskge_b200.so not found at %s: run `python scikit-kge_b200/build.py` '
                '(there is no CPU fallback)' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib():
    """Library handle for compute calls: also requires an sm_100 device."""
    global _device_ok
    l = load_library()
    if not _device_ok:
        if not torch.cuda.is_available():
            raise RuntimeError('skge (B200 build) needs a CUDA device; there is no CPU fallback')
        torch.cuda.current_device()
        torch.zeros(1, device='cuda')  # make sure the primary context exists
        rc = l.skge_check_device()
        if rc != 0:
            raise RuntimeError(l.skge_last_error().decode())
        _device_ok = True
    return l


def check(rc):
    if rc != 0:
        msg = load_library().skge_last_error().decode()
        if rc == -10001:
            raise ValueError(msg)
        raise RuntimeError('libskge_b200 error %d: %s' % (rc, msg))


def ptr(t):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('skge (B200 build) computes on CUDA tensors only; there is no CPU fallback')
    assert t.is_contiguous(), 'expected a contiguous CUDA tensor'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def device():
    """Where parameters live.  Without a CUDA device tensors can still be *held* (host
    storage for pickling / inspection); every compute call goes through lib(), which raises."""
    if not torch.cuda.is_available():
        return torch.device('cpu')
    return torch.device('cuda', torch.cuda.current_device())


class Workspace(object):
    """Grow-only scratch buffer handed to the library (it never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes):
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device():
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device())
        return self.buf


def as_i32(x):
    """Host list / ndarray / tensor -> contiguous int32 CUDA tensor."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device(), dtype=torch.int32).contiguous()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.int32)).to(device())


def as_f32(x):
    if isinstance(x, torch.Tensor):
        return x.to(device=device(), dtype=torch.float32).contiguous()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32)).to(device())
