import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from speech_recognition_hmm_continuous_b200 import api, synth
from oracle import oracle as o
# decode through k_emis_ws<false> (M > 16) and training emissions through k_emis_ws<true>
for (V, N, M, U, seed) in ((3, 3, 128, 4, 21), (3, 5, 32, 6, 22), (2, 2, 160, 4, 23)):
    cen, s = synth.make_centres(V, N, M, 39, seed=seed)
    labels = np.arange(U) % V
    x, off = synth.make_utterances(cen, s, labels, seed=seed + 1, tmin=40, tmax=70)
    ms = api.ModelSet.from_dict(synth.make_models(cen, s))
    want = np.array([[o.forward_score(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v]), x[off[u]:off[u + 1]]) for v in range(V)] for u in range(U)])
    res = {}
    for f16 in (0, 1):
        ctx = api.Context(0)
        ctx.set_option("dec_f16", f16)
        ctx.set_features(x, off); ctx.set_models(ms)
        res[f16] = ctx.forward_scores()
        ctx.close()
    print("decode V%d N%d M%d: rel err tf32 %.3e  f16 %.3e" % (V, N, M, np.abs(res[0] / want - 1).max(), np.abs(res[1] / want - 1).max()))
for (V, N, M, U, seed) in ((3, 5, 16, 9, 31), (2, 5, 3, 6, 32), (1, 3, 128, 2, 33)):
    cen, s = synth.make_centres(V, N, M, 39, seed=seed)
    labels = np.arange(U) % V
    x, off = synth.make_utterances(cen, s, labels, seed=seed + 1, tmin=40, tmax=70)
    ms = api.ModelSet.from_dict(synth.make_models(cen, s))
    res = {}
    for f16 in (0, 1):
        ctx = api.Context(0)
        ctx.set_option("train_f16", f16)
        ctx.set_features(x, off); ctx.set_models(ms)
        lb, post = ctx.emissions(1, labels[1])
        st, lpu = ctx.estep(labels)
        res[f16] = (lb, lpu, st)
        ctx.close()
    b, p = o.emissions(o.Model(ms.A[labels[1]], ms.c[labels[1]], ms.mu[labels[1]], ms.iv[labels[1]], ms.det[labels[1]]), x[off[1]:off[2]])
    print("train V%d N%d M%d: logP rel diff f16 vs tf32 %.3e ; stats max rel diff %.3e" % (V, N, M, np.abs(res[1][1] / res[0][1] - 1).max(),
          np.abs(res[1][2] - res[0][2]).max() / np.abs(res[0][2]).max()))
