"""A few EM iterations (and one decode pass) at the C2 shape, for ncu captures."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 10, 5, 16, 1000
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels = (np.arange(U) % V).astype(np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=1234)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_features(x, off); ctx.set_models(ms); ctx.em_reset()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.estep(labels, download=False, want_logp=False)
    ctx.mstep(threshold=-1.0)
if len(sys.argv) > 2:
    ctx.rank(ctx.forward_scores())
    ctx.viterbi(labels)
ctx.synchronize()
ctx.close()
