"""Second-order sums of k_accum_h / k_accum_ws against the oracle on the (5, 16, 2, 6) case of test_estep_statistics_match_oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as o
from speech_recognition_hmm_continuous_b200 import api, synth
N, M, V, U = 5, 16, 2, 6
seed = 31 + N + M
cen, s = synth.make_centres(V, N, M, 39, seed=seed)
labels = np.arange(U) % V
x, off = synth.make_utterances(cen, s, labels, seed=seed + 1, tmin=60, tmax=120)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_features(x, off)
ctx.set_models(ms)
for h in (0, 1):
    ctx.set_option("h_acc", h)
    stats, lpu = ctx.estep(labels)
    for v in range(V):
        us = np.nonzero(labels == v)[0]
        xv = np.concatenate([x[off[u]:off[u + 1]] for u in us])
        offv = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])])
        st, lp = o.estep(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v], ms.words[v]), xv, offv)
        sp = api.split_stats(stats[v], N, M, 39)
        S0 = np.maximum(st.S0, 1e-300)[..., None]
        occ = st.S0 > 1e-3 * st.S0.max()
        want2 = st.S2c / S0
        scale2 = want2 + (st.S1 / S0 - x.mean(0)) ** 2
        err = np.abs(sp["S2c"] / S0 - want2)
        tol = 1e-4 * want2 + 2e-6 * scale2 + 1e-8
        ratio = np.where(occ[..., None] if occ.ndim < err.ndim else occ, err / tol, 0)
        i = np.unravel_index(np.argmax(ratio), ratio.shape)
        print("h=%d v=%d worst err/tol %.3f at %s: err %.3e want2 %.3e scale2 %.3e S0 %.4g | S0 rel err max %.2e, S1 err/sd max %.2e" % (
            h, v, ratio.max(), i, err[i], want2[i], scale2[i], st.S0[i[:-1]] if st.S0.ndim == len(i) - 1 else st.S0.reshape(-1)[i[0]],
            (np.abs(sp["S0"] - st.S0) / st.S0.max()).max(),
            (np.abs(sp["S1"] / S0 - st.S1 / S0) / np.maximum(np.abs(st.S1 / S0), np.sqrt(want2)))[occ].max()))
        print("   fraction above tol: %.4f; radius per dim max %.3f, sd typical %.3f" % ((ratio > 1).mean(), np.abs(x - x.mean(0)).max(), np.sqrt(np.median(want2))))
