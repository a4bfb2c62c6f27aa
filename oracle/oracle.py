"""ctypes/numpy front end of oracle/liboracle.so (hmm_oracle.c).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def build():
    src = os.path.join(_HERE, "hmm_oracle.c")
    out = os.path.join(_HERE, "liboracle.so")
    if (not os.path.exists(out)) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src, "-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_gauss.restype = C.c_double
        _LIB.orc_logprob.restype = C.c_double
        _LIB.orc_estep.restype = C.c_double
        _LIB.orc_forward_score.restype = C.c_double
        _LIB.orc_viterbi.restype = C.c_double
        _LIB.orc_train.restype = C.c_int
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """One word model in the reference's semantics: iv = INVERSE variance,
    det = product of variances."""

    def __init__(self, A, c, mu, iv, det, word="w"):
        self.A = _f64(A).copy()
        self.c = _f64(c).copy()
        self.mu = _f64(mu).copy()
        self.iv = _f64(iv).copy()
        self.det = _f64(det).copy()
        self.word = word
        self.N, self.M, self.D = self.mu.shape
        assert self.A.shape == (self.N, self.N) and self.c.shape == (self.N, self.M)
        assert self.iv.shape == self.mu.shape and self.det.shape == (self.N, self.M)

    def copy(self):
        return Model(self.A, self.c, self.mu, self.iv, self.det, self.word)


def gauss(x, mu, iv, det):
    x, mu, iv = _f64(x), _f64(mu), _f64(iv)
    return lib().orc_gauss(len(x), _d(x), _d(mu), _d(iv), C.c_double(det))


def emissions(m, x, want_post=True):
    x = _f64(x)
    T = x.shape[0]
    b = np.zeros((T, m.N))
    post = np.zeros((T, m.N, m.M)) if want_post else None
    lib().orc_emissions(m.N, m.M, m.D, T, _d(x), _d(m.c), _d(m.mu), _d(m.iv), _d(m.det), _d(b),
                        _d(post) if want_post else None)
    return b, post


def forward(m, b):
    b = _f64(b)
    T = b.shape[0]
    alpha = np.zeros((T, m.N))
    scale = np.zeros(T)
    lib().orc_forward(m.N, T, _d(m.A), _d(b), _d(alpha), _d(scale))
    return alpha, scale


def backward(m, b, scale):
    b, scale = _f64(b), _f64(scale)
    T = b.shape[0]
    beta = np.zeros((T, m.N))
    lib().orc_backward(m.N, T, _d(m.A), _d(b), _d(scale), _d(beta))
    return beta


def logprob(alpha, scale):
    alpha, scale = _f64(alpha), _f64(scale)
    T, N = alpha.shape
    return lib().orc_logprob(N, T, _d(alpha), _d(scale))


class Stats:
    def __init__(self, N, M, D):
        self.num_trans = np.zeros((N, N))
        self.den_trans = np.zeros(N)
        self.den_mix = np.zeros(N)
        self.S0 = np.zeros((N, M))
        self.S1 = np.zeros((N, M, D))
        self.S2c = np.zeros((N, M, D))
        self.sum_logp = 0.0
        self.n_utt = 0


def estep(m, x, off):
    """E-step of one model over utterances x[off[u]:off[u+1]].  Returns (Stats, logp_per_utt)."""
    x = _f64(x)
    off = np.ascontiguousarray(off, dtype=np.int64)
    U = len(off) - 1
    st = Stats(m.N, m.M, m.D)
    lpu = np.zeros(U)
    st.sum_logp = lib().orc_estep(m.N, m.M, m.D, U, off.ctypes.data_as(_lp), _d(x), _d(m.A), _d(m.c),
                                  _d(m.mu), _d(m.iv), _d(m.det), _d(st.num_trans), _d(st.den_trans),
                                  _d(st.den_mix), _d(st.S0), _d(st.S1), _d(st.S2c), _d(lpu))
    st.n_utt = U
    return st, lpu


def mstep(m, st):
    """In-place M-step (A10)."""
    lib().orc_mstep(m.N, m.M, m.D, _d(st.num_trans), _d(st.den_trans), _d(st.den_mix), _d(st.S0),
                    _d(st.S1), _d(st.S2c), _d(m.A), _d(m.c), _d(m.mu), _d(m.iv), _d(m.det))
    return m


def train(m, x, off, max_iter=0):
    """EM loop (A11) in place.  Returns (iterations, mean logP of the last E-step)."""
    x = _f64(x)
    off = np.ascontiguousarray(off, dtype=np.int64)
    mean = C.c_double(0.0)
    it = lib().orc_train(m.N, m.M, m.D, len(off) - 1, off.ctypes.data_as(_lp), _d(x), _d(m.A), _d(m.c),
                         _d(m.mu), _d(m.iv), _d(m.det), C.byref(mean), int(max_iter))
    return it, mean.value


def init_model(N, M, x, off, word="w"):
    x = _f64(x)
    D = x.shape[1]
    off = np.ascontiguousarray(off, dtype=np.int64)
    m = Model(np.zeros((N, N)), np.zeros((N, M)), np.zeros((N, M, D)), np.zeros((N, M, D)),
              np.zeros((N, M)), word)
    lib().orc_init_model(N, M, D, len(off) - 1, off.ctypes.data_as(_lp), _d(x), _d(m.A), _d(m.c),
                         _d(m.mu), _d(m.iv), _d(m.det))
    return m


def forward_score(m, x):
    x = _f64(x)
    return lib().orc_forward_score(m.N, m.M, m.D, x.shape[0], _d(x), _d(m.A), _d(m.c), _d(m.mu),
                                   _d(m.iv), _d(m.det))


def rank(score):
    score = _f64(score)
    idx = np.zeros(len(score), dtype=np.int32)
    lib().orc_rank(len(score), _d(score), idx.ctypes.data_as(_ip))
    return idx


def viterbi(m, b):
    """b = LINEAR emission densities [T][N] (as orc_emissions returns).  -> (score, path int32[T])."""
    b = _f64(b)
    T = b.shape[0]
    path = np.zeros(T, dtype=np.int32)
    with np.errstate(divide="ignore"):
        s = lib().orc_viterbi(m.N, T, _d(m.A), _d(b), path.ctypes.data_as(_ip))
    return s, path
