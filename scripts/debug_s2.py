"""Where does the tensor-core accumulate path lose accuracy in S2c?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
from oracle import oracle as o
for (N, M, V, U) in [(5, 16, 2, 6), (3, 128, 1, 2)]:
    seed = 31 + N + M
    cen, s = synth.make_centres(V, N, M, 39, seed=seed)
    labels = np.arange(U) % V
    x, off = synth.make_utterances(cen, s, labels, seed=seed + 1, tmin=60, tmax=120)
    ms = api.ModelSet.from_dict(synth.make_models(cen, s))
    res = {}
    for path in (0, 2):
        ctx = api.Context(0)
        ctx.set_option("tc_emis", path)
        ctx.set_features(x, off); ctx.set_models(ms)
        stats, lpu = ctx.estep(labels)
        res[path] = stats
        ctx.close()
    for v in range(V):
        us = np.nonzero(labels == v)[0]
        xv = np.concatenate([x[off[u]:off[u + 1]] for u in us])
        offv = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])])
        st, lp = o.estep(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v]), xv, offv)
        S0 = np.maximum(st.S0, 1e-300)[..., None]
        occ = st.S0 > 1e-3 * st.S0.max()
        want = (st.S2c / S0)
        for path in (0, 2):
            sp = api.split_stats(res[path][v], N, M, 39)
            got = sp["S2c"] / S0
            rel = np.abs(got - want) / np.abs(want)
            rel[~occ] = 0
            i = np.unravel_index(np.argmax(rel), rel.shape)
            rs0 = np.abs(sp["S0"] - st.S0) / np.maximum(st.S0, 1e-300); rs0[~occ] = 0
            print("N%d M%d v%d path %d: S2c/S0 max rel err %.3g at %s (S0=%.4g, want %.6g got %.6g, mu-centre/sd %.3g); S0 max rel err %.3g"
                  % (N, M, v, path, rel.max(), i, st.S0[i[0], i[1]], want[i], got[i], 0.0, rs0.max()))
            # error budget: raw pieces
            if path == 2:
                cnt = (rel > 1e-4).sum()
                print("   elements over 1e-4:", cnt, "of", occ.sum() * 39)
