"""Second, independent pins for the Viterbi decoder (TEST INFRASTRUCTURE).

The reference has no Viterbi (R-FS:10-11, 1024: "Algorithm used for recognition: Forward"), so the C
restatement `orc_viterbi` (oracle/hmm_oracle.c) shares an author with the CUDA kernel.  These two
restatements are written from the reference's conventions alone and share no code with either:

  np_log_emissions  log b_i(t) from calc_gaus / calc_symbol_probab (T-FS:1804-1841, 1749-1783) in log form
  np_viterbi        vectorised log-domain dynamic programme: pi = [1,0,..] (T-FS:232-234), the full A
                    matrix as calc_alpha uses it (T-FS:1420-1442), termination in the final state as
                    calc_beta / calc_probability do (T-FS:1484-1490, 1546-1549); numpy.argmax returns the
                    first maximum, i.e. the lowest predecessor wins a tie (R-FS:984, strict compare)
  brute_viterbi     every one of the N^T state sequences, for tiny cases (N <= 3, T <= 8): the best
                    sequence that starts in state 0 and ends in state N-1; among equal scores the one that
                    is smallest when compared from the LAST frame backwards (what a back-trace that takes
                    the lowest predecessor produces)
"""
import itertools

import numpy as np


def np_log_emissions(c, mu, iv, det, x):
    """c[N][M], mu[N][M][D], iv[N][M][D] (inverse variances), det[N][M] (product of variances), x[T][D] -> logb[T][N]."""
    x = np.asarray(x, dtype=np.float64)
    D = x.shape[1]
    dif = x[:, None, None, :] - mu[None]                                   # [T][N][M][D]
    q = -0.5 * np.einsum("tnmd,nmd,tnmd->tnm", dif, iv, dif)
    with np.errstate(divide="ignore"):
        lg = q - 0.5 * (D * np.log(2.0 * np.pi) + np.log(np.abs(det)))[None] + np.log(c)[None]
    mx = lg.max(axis=2, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    with np.errstate(divide="ignore"):
        return (mx + np.log(np.exp(lg - mx).sum(axis=2, keepdims=True)))[..., 0]


def np_viterbi(A, logb):
    """-> (score, path int32[T]); score = -inf and an all-zero path when the final state is unreachable."""
    T, N = logb.shape
    with np.errstate(divide="ignore"):
        la = np.log(np.asarray(A, dtype=np.float64))
    delta = np.full(N, -np.inf)
    delta[0] = 0.0
    delta = delta + logb[0]
    psi = np.zeros((T, N), dtype=np.int32)
    for t in range(1, T):
        cand = delta[:, None] + la                                           # [from][to]
        psi[t] = np.argmax(cand, axis=0)
        delta = cand[psi[t], np.arange(N)] + logb[t]
    path = np.zeros(T, dtype=np.int32)
    s = N - 1
    for t in range(T - 1, -1, -1):
        path[t] = s
        s = psi[t, s]
    return float(delta[N - 1]), path


def brute_viterbi(A, logb):
    T, N = logb.shape
    assert N ** T <= 20000, "tiny cases only"
    with np.errstate(divide="ignore"):
        la = np.log(np.asarray(A, dtype=np.float64))
    best, arg = -np.inf, None
    for seq in itertools.product(range(N), repeat=T):
        if seq[0] != 0 or seq[-1] != N - 1:
            continue
        s = 0.0 + logb[0, 0]
        for t in range(1, T):
            s = (s + la[seq[t - 1], seq[t]]) + logb[t, seq[t]]              # the dynamic programme's association
        key = tuple(reversed(seq))
        if s > best or (s == best and arg is not None and key < tuple(reversed(arg))):
            best, arg = s, seq
    if arg is None:
        return -np.inf, np.zeros(T, dtype=np.int32)
    return float(best), np.array(arg, dtype=np.int32)
