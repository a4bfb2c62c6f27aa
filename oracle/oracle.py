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


# ---- multi-stream models (param_number > 1) ---------------------------------------------------
# The reference multiplies the per-stream mixture densities inside calc_alpha / calc_beta /
# calc_transition_probab (T-FS:1406-1409, 1427-1431, 1500-1503, 1606-1609) and keeps one set of mixture
# accumulators per stream (T-FS:303-318); the recogniser does the same product (R-FS:341-364, 762-791).
# Restated here on top of the single-stream primitives above (which are pinned against the compiled
# reference): the product of the streams' b is handed to orc_forward / orc_backward, the accumulators
# follow calc_transition_probab / calc_den_mix_coef / calc_mix_param (T-FS:1577-1727) in numpy.
# Pinned against the reference's own trainer and recogniser run with two streams: tests/golden/synth_p2.npz.
def estep_streams(models, xs, off, delta=1):
    """models[p], xs[p] = model / features of stream p (same A, same utterance lengths).
    -> ([Stats per stream] (transition statistics identical in all), logp per utterance)."""
    P = len(models)
    m0 = models[0]
    N = m0.N
    off = np.ascontiguousarray(off, dtype=np.int64)
    U = len(off) - 1
    sts = [Stats(m.N, m.M, m.D) for m in models]
    lpu = np.zeros(U)
    for u in range(U):
        sl = slice(int(off[u]), int(off[u + 1]))
        bp = [emissions(models[p], xs[p][sl]) for p in range(P)]
        b = bp[0][0].copy()
        for p in range(1, P):
            b = b * bp[p][0]                                  # product *= symbol_probab[j][i][k], stream order
        alpha, scale = forward(m0, b)
        beta = backward(m0, b, scale)
        lpu[u] = logprob(alpha, scale)
        T = b.shape[0]
        num = np.zeros((N, N)); den = np.zeros(N)
        for i in range(N):                                    # calc_transition_probab, T-FS:1577-1620
            for j in range(i, min(N, i + delta + 1)):
                num[i, j] = np.sum(alpha[:T - 1, i] * m0.A[i, j] * b[1:, j] * beta[1:, j])
            den[i] = np.sum(alpha[:T - 1, i] * beta[:T - 1, i] / scale[:T - 1])
        gamma = alpha * beta / scale[:, None]                 # calc_den_mix_coef / calc_mix_param
        for p in range(P):
            st, m, x = sts[p], models[p], _f64(xs[p])[sl]
            st.num_trans += num; st.den_trans += den; st.den_mix += gamma.sum(0)
            w = gamma[:, :, None] * bp[p][1]                  # [T][N][M]
            st.S0 += w.sum(0)
            st.S1 += np.einsum("tnm,td->nmd", w, x)
            dev = x[:, None, None, :] - m.mu[None]
            st.S2c += np.einsum("tnm,tnmd->nmd", w, dev * dev)
            st.sum_logp += lpu[u]; st.n_utt += 1
    return sts, lpu


def train_streams(models, xs, off, max_iter=0, threshold=1e-3):
    """EM loop of T-FS:238-361 for one word with several streams, in place.  -> (iterations, mean logP)."""
    old, it = 1.0, 0
    while True:
        it += 1
        sts, lpu = estep_streams(models, xs, off)
        probab = float(np.sum(lpu))
        var = abs((old - probab) / old)
        if not var > threshold or (max_iter and it >= max_iter):
            return it, probab / len(lpu)
        old = probab
        A = None
        for m, st in zip(models, sts):
            mstep(m, st)
            A = m.A if A is None else A
            m.A[...] = A                                      # one transition matrix (identical statistics anyway)


def forward_score_streams(models, xs):
    b = None
    for m, x in zip(models, xs):
        bp = emissions(m, x, want_post=False)[0]
        b = bp if b is None else b * bp
    alpha, scale = forward(models[0], b)
    with np.errstate(divide="ignore", invalid="ignore"):
        return logprob(alpha, scale)
