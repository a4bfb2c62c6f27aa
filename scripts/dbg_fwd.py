import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from speech_recognition_hmm_continuous_b200 import api, synth
from oracle import oracle as o
cen, s = synth.make_centres(7, 5, 3, 39, seed=7)
labels = np.arange(21) % 7
x, off = synth.make_utterances(cen, s, labels, seed=8, tmin=60, tmax=120)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_features(x, off); ctx.set_models(ms)
got = ctx.forward_scores()
ctx.set_option("fwd_f64", 1)
got64 = ctx.forward_scores()
want = np.array([[o.forward_score(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v], ms.words[v]), x[off[u]:off[u + 1]]) for v in range(ms.V)] for u in range(21)])
np.set_printoptions(linewidth=200, precision=6)
print("rel32", np.abs(got / want - 1).max(), "rel64", np.abs(got64 / want - 1).max())
r = np.abs(got / want - 1)
i = np.unravel_index(np.argmax(r), r.shape)
print(i, got[i], got64[i], want[i])
print(got[:3]); print(want[:3])
