"""CPU: the C oracle's Viterbi against the two independent restatements of tests/viterbi_pins.py
(brute-force enumeration of every state sequence, and a numpy log-domain dynamic programme)."""
import numpy as np
import pytest

from oracle import oracle as o
from viterbi_pins import brute_viterbi, np_log_emissions, np_viterbi


def _random_case(rng, N, T, full):
    A = rng.uniform(0.05, 1.0, size=(N, N))
    if not full:
        A = np.triu(A) - np.triu(A, 2)          # the reference's DELTA = 1 band (T-FS:774-795)
    A /= A.sum(axis=1, keepdims=True)
    b = np.exp(rng.uniform(-30.0, 0.0, size=(T, N)))
    return A, b


@pytest.mark.parametrize("N,T,full", [(1, 1, False), (1, 5, False), (2, 2, False), (2, 8, True), (3, 3, False), (3, 8, False), (3, 8, True), (3, 2, False)])
def test_oracle_viterbi_equals_brute_force_and_numpy(N, T, full):
    rng = np.random.default_rng(1000 + 17 * N + T + (5 if full else 0))
    for _ in range(25):
        A, b = _random_case(rng, N, T, full)
        m = o.Model(A, np.ones((N, 1)), np.zeros((N, 1, 1)), np.ones((N, 1, 1)), np.ones((N, 1)))
        s_o, p_o = o.viterbi(m, b)
        s_n, p_n = np_viterbi(A, np.log(b))
        s_b, p_b = brute_viterbi(A, np.log(b))
        if not np.isfinite(s_b):            # T < N with the band: the final state cannot be reached
            assert s_o == -np.inf and s_n == -np.inf
            continue
        assert s_o == s_n == s_b            # the same association of the same doubles
        assert (p_o == p_n).all() and (p_o == p_b).all()


def test_ties_go_to_the_lowest_predecessor():
    """Exact ties (equal emissions, equal transition entries): all three restatements take the lowest index."""
    A = np.array([[0.25, 0.25, 0.5], [0.0, 0.5, 0.5], [0.0, 0.0, 1.0]])
    for T in (3, 4, 6, 8):
        b = np.full((T, 3), 0.125)
        m = o.Model(A, np.ones((3, 1)), np.zeros((3, 1, 1)), np.ones((3, 1, 1)), np.ones((3, 1)))
        s_o, p_o = o.viterbi(m, b)
        s_n, p_n = np_viterbi(A, np.log(b))
        s_b, p_b = brute_viterbi(A, np.log(b))
        assert s_o == s_n == s_b
        assert (p_o == p_n).all() and (p_o == p_b).all(), (T, p_o, p_n, p_b)


def test_numpy_log_emissions_equal_the_oracle():
    rng = np.random.default_rng(7)
    N, M, D, T = 3, 4, 6, 20
    c = rng.uniform(0.1, 1.0, size=(N, M)); c /= c.sum(axis=1, keepdims=True)
    mu = rng.standard_normal((N, M, D))
    var = rng.uniform(0.5, 2.0, size=(N, M, D))
    x = rng.standard_normal((T, D))
    A = np.triu(np.ones((N, N))) - np.triu(np.ones((N, N)), 2); A /= A.sum(axis=1, keepdims=True)
    m = o.Model(A, c, mu, 1.0 / var, var.prod(axis=2))
    b, _ = o.emissions(m, x, want_post=False)
    assert np.allclose(np.log(b), np_log_emissions(c, mu, 1.0 / var, var.prod(axis=2), x), rtol=1e-12, atol=1e-12)
