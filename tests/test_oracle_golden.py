"""The CPU oracle (oracle/hmm_oracle.c) against the committed golden vectors that the REFERENCE
produced (tests/golden/make_golden.py).  Runs anywhere: needs neither /root/reference nor a GPU."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as o
from oracle import ref as r
from speech_recognition_hmm_continuous_b200 import synth


def _max_rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def test_shipped_kats_train(golden_dir):
    """13 shipped feature files, N=6 M=1: mean logP to the 6 printed decimals and iteration count
    (SURVEY.md section 4.5 table); trained models bit-exact with what the reference wrote."""
    kat = json.load(open(os.path.join(golden_dir, "kat_diag.json")))
    km = np.load(os.path.join(golden_dir, "kat_models.npz"))
    words = [str(w) for w in km["words"]]
    assert len(kat["train"]) == 13
    for word, rec in kat["train"].items():
        x = r.read_features(os.path.join(golden_dir, "perfil", rec["file"]))
        off = np.array([0, len(x)])
        m = o.init_model(6, 1, x, off, word)
        its, mean = o.train(m, x, off)
        assert its == rec["iterations"], word
        assert "%.6f" % mean == "%.6f" % rec["mean_logp"], word
        k = words.index(word)
        for name in ("A", "c", "mu", "iv", "det"):
            assert _max_rel(getattr(m, name), km[name][k]) < 1e-12, (word, name)


def test_shipped_kats_recognition_degenerate(golden_dir):
    """Recogniser on the 13 diagonal models: the NaN / -inf regime.  The oracle's score strings and
    full sorted order must equal what the reference printed (R2/R3 incl. NaN barriers)."""
    kat = json.load(open(os.path.join(golden_dir, "kat_diag.json")))
    km = np.load(os.path.join(golden_dir, "kat_models.npz"))
    words = [str(w) for w in km["words"]]
    models = [o.Model(km["A"][k], km["c"][k], km["mu"][k], km["iv"][k], km["det"][k], words[k]) for k in range(13)]
    files = sorted(glob.glob(os.path.join(golden_dir, "perfil", "*.perfil")))
    assert len(files) == 13
    ncorrect = 0
    for rec, f in zip(kat["recognition"], files):
        x = r.read_features(f)
        with np.errstate(all="ignore"):
            score = np.array([o.forward_score(m, x) for m in models])
        idx = o.rank(score)
        got = [[words[i], "%f" % score[i]] for i in idx]
        want = [[w, "nan" if "nan" in v else v] for w, v in rec["sorted"]]  # C prints "-nan", Python "nan"
        assert got == want, rec["spoken"]
        ncorrect += words[idx[0]] == rec["spoken"]
    assert ncorrect == 2  # "Correct words: 2" in the reference's result file
    assert "Correct words: 2\nErrors: 11" in kat["result_file"]


def _synth_c1(g):
    V, N, M, D = int(g["V"]), int(g["N"]), int(g["M"]), int(g["D"])
    cen, s = synth.make_centres(V, N, M, D, seed=1234)
    x, off = synth.make_utterances(cen, s, g["train_labels"], seed=1234, tmin=70, tmax=110)
    xt, offt = synth.make_utterances(cen, s, g["test_labels"], seed=4321, tmin=70, tmax=110)
    return V, N, M, D, x, off, xt, offt


def test_synth_c1_train_and_recognise(golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_c1.npz"))
    V, N, M, D, x, off, xt, offt = _synth_c1(g)
    models = []
    for v in range(V):
        us = np.nonzero(g["train_labels"] == v)[0]
        xv = np.concatenate([x[off[u]:off[u + 1]] for u in us])
        offv = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])])
        m = o.init_model(N, M, xv, offv)
        for name in ("A", "c", "mu", "iv", "det"):
            assert _max_rel(getattr(m, name), g["init_" + name][v]) < 1e-13, (v, name)
        st, lpu = o.estep(m, xv, offv)
        for name in ("num_trans", "den_trans", "den_mix", "S0", "S1", "S2c"):
            assert _max_rel(getattr(st, name), g["stat_" + name][v]) < 1e-12, (v, name)
        assert abs(st.sum_logp - g["stat_sum_logp"][v]) < 1e-9 * abs(st.sum_logp)
        its, mean = o.train(m, xv, offv)
        assert its == g["iterations"][v]
        assert "%.6f" % mean == "%.6f" % g["mean_logp"][v]
        for name in ("A", "c", "mu", "iv", "det"):
            assert _max_rel(getattr(m, name), g["trained_" + name][v]) < 1e-12, (v, name)
        models.append(m)
    for u in range(len(offt) - 1):
        sc = np.array([o.forward_score(m, xt[offt[u]:offt[u + 1]]) for m in models])
        assert np.allclose(sc, g["score"][u], rtol=1e-13, atol=0)
        assert (o.rank(sc) == g["order"][u]).all()
        assert o.rank(sc)[0] == g["test_labels"][u]


def test_rank_rule_nan_barriers():
    """R3: a NaN never moves; each NaN-free run is sorted on its own; -inf compares normally;
    ties keep the lower index first."""
    sc = np.array([1.0, np.nan, 3.0, -np.inf, 3.0, 5.0, np.nan, 7.0])
    assert list(o.rank(sc)) == [0, 1, 5, 2, 4, 3, 6, 7]
    assert list(o.rank(np.array([2.0, 2.0, 2.0]))) == [0, 1, 2]
    assert o.rank(np.array([np.nan, 5.0, 9.0]))[0] == 0


def test_viterbi_properties():
    """Viterbi is absent from the reference (parity unpinned); pin it by properties instead:
    score <= forward logP, equality when N == 1, path is a valid left-to-right walk ending in N-1,
    and the score equals the path's own log-probability."""
    cen, s = synth.make_centres(2, 5, 3, 39, seed=5)
    mods = synth.make_models(cen, s)
    x, off = synth.make_utterances(cen, s, [0, 1, 0], seed=6, tmin=50, tmax=80)
    for u in range(3):
        for v in range(2):
            m = o.Model(mods["A"][v], mods["c"][v], mods["mu"][v], mods["iv"][v], mods["det"][v])
            xu = x[off[u]:off[u + 1]]
            b, _ = o.emissions(m, xu, want_post=False)
            sc, path = o.viterbi(m, b)
            assert sc <= o.forward_score(m, xu) + 1e-9
            assert path[0] == 0 and path[-1] == m.N - 1
            assert ((np.diff(path) == 0) | (np.diff(path) == 1)).all()
            with np.errstate(divide="ignore"):
                lp = np.log(b[0, 0]) + sum(np.log(m.A[path[t - 1], path[t]]) + np.log(b[t, path[t]]) for t in range(1, len(path)))
            assert abs(lp - sc) < 1e-9 * abs(sc)
    m1 = o.Model(np.ones((1, 1)), mods["c"][0][:1], mods["mu"][0][:1], mods["iv"][0][:1], mods["det"][0][:1])
    b, _ = o.emissions(m1, x[: off[1]], want_post=False)
    sc, path = o.viterbi(m1, b)
    assert abs(sc - o.forward_score(m1, x[: off[1]])) < 1e-9 * abs(sc)


# ---- two feature streams (param_number = 2): the reference's own trainer and recogniser ----
def _p2_word(g, v, train=True):
    lab = g["train_labels"] if train else g["test_labels"]
    off = g["off"] if train else g["offt"]
    us = np.nonzero(lab == v)[0] if train else np.arange(len(lab))
    xs = []
    for p in range(2):
        x = g[("x%d" if train else "xt%d") % p]
        xs.append(np.concatenate([x[off[u]:off[u + 1]] for u in us]))
    o2 = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])]).astype(np.int64)
    return xs, o2


def test_two_stream_training_matches_reference(golden_dir):
    """Multi-stream restatement (oracle.estep_streams / train_streams) against the reference trainer run with
    param_number = 2: iterations, mean log-probability and every trained parameter of both streams."""
    g = np.load(os.path.join(golden_dir, "synth_p2.npz"))
    N = int(g["N"])
    for v in range(int(g["V"])):
        xs, off = _p2_word(g, v)
        models = [o.init_model(N, int(g["M"][p]), xs[p], off) for p in range(2)]
        assert np.array_equal(models[0].A, models[1].A)
        its, mean = o.train_streams(models, xs, off)
        assert its == g["iterations"][v]
        assert abs(mean - g["mean_logp"][v]) <= 1e-9 * abs(g["mean_logp"][v]) + 1e-6   # report prints %f
        for p in range(2):
            for k in ("A", "c", "mu", "iv", "det"):
                want = g["trained_s%d_%s" % (p, k)][v]
                assert np.allclose(getattr(models[p], k), want, rtol=1e-9, atol=1e-12), (v, p, k)


def test_two_stream_recognition_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_p2.npz"))
    xs, off = _p2_word(g, 0, train=False)
    for u in range(len(off) - 1):
        for v in range(int(g["V"])):
            ms = [o.Model(*[g["trained_s%d_%s" % (p, k)][v] for k in ("A", "c", "mu", "iv", "det")]) for p in range(2)]
            s = o.forward_score_streams(ms, [x[off[u]:off[u + 1]] for x in xs])
            assert abs(s - g["score"][u, v]) <= 1e-6 + 1e-9 * abs(s)                       # printed with %f
    assert (np.argmax(g["score"], axis=1) == g["test_labels"]).all()


def test_multi_stream_restatement_reduces_to_the_single_stream_oracle():
    """oracle.estep_streams (numpy, on top of the pinned primitives) with ONE stream is the C restatement's E-step."""
    cen, s = synth.make_centres(1, 4, 3, 7, seed=3)
    x, off = synth.make_utterances(cen, s, [0] * 5, seed=4, tmin=30, tmax=50)
    d = synth.make_models(cen, s)
    m = o.Model(d["A"][0], d["c"][0], d["mu"][0], d["iv"][0], d["det"][0])
    want, lp = o.estep(m, x, off)
    (got,), lps = o.estep_streams([m], [x], off)
    assert np.allclose(lps, lp, rtol=1e-13)
    for k in ("num_trans", "den_trans", "den_mix", "S0", "S1", "S2c"):
        assert np.allclose(getattr(got, k), getattr(want, k), rtol=1e-10, atol=1e-12 * np.abs(getattr(want, k)).max()), k
    assert abs(got.sum_logp - want.sum_logp) <= 1e-12 * abs(want.sum_logp) and got.n_utt == want.n_utt
    # the order of the streams does not matter to the product (up to rounding)
    cen2, s2 = synth.make_centres(1, 4, 2, 5, seed=8)
    x2, off2 = synth.make_utterances(cen2, s2, [0] * 5, seed=4, tmin=30, tmax=50)
    assert np.array_equal(off, off2)
    d2 = synth.make_models(cen2, s2)
    m2 = o.Model(d["A"][0], d2["c"][0], d2["mu"][0], d2["iv"][0], d2["det"][0])
    (a1, a2), lpa = o.estep_streams([m, m2], [x, x2], off)
    (b2, b1), lpb = o.estep_streams([m2, m], [x2, x], off)
    assert np.allclose(lpa, lpb, rtol=1e-12) and np.allclose(a1.S1, b1.S1, rtol=1e-9, atol=1e-9) and np.allclose(a2.S0, b2.S0, rtol=1e-9, atol=1e-12)
