import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from speech_recognition_hmm_continuous_b200 import api, synth
from oracle import oracle as o
for (V, N, M, U, seed) in ((7, 5, 3, 14, 7), (4, 5, 16, 8, 11), (3, 3, 8, 6, 13)):
    cen, s = synth.make_centres(V, N, M, 39, seed=seed)
    labels = np.arange(U) % V
    x, off = synth.make_utterances(cen, s, labels, seed=seed + 1, tmin=60, tmax=120)
    ms = api.ModelSet.from_dict(synth.make_models(cen, s))
    want = np.array([[o.forward_score(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v]), x[off[u]:off[u + 1]]) for v in range(V)] for u in range(U)])
    res = {}
    for f16 in (0, 1):
        ctx = api.Context(0)
        ctx.set_option("dec_f16", f16)
        ctx.set_features(x, off); ctx.set_models(ms)
        res[f16] = ctx.forward_scores()
        ctx.close()
    print("V%d N%d M%d: rel err tf32 %.3e  f16 %.3e  | f16 vs tf32 %.3e | abs max f16 %.3e tf32 %.3e" % (
        V, N, M, np.abs(res[0] / want - 1).max(), np.abs(res[1] / want - 1).max(), np.abs(res[1] / res[0] - 1).max(),
        np.abs(res[1] - want).max(), np.abs(res[0] - want).max()))
