"""Decode legs at a given shape with wall-clock prints (debugging aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = [int(a) for a in sys.argv[1:5]]
cen, s = synth.make_centres(V, N, M, 39, seed=77)
labels = (np.arange(U) % V).astype(np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=78)
ctx = api.Context(0, timing=True)
ctx.set_features(x, off); ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
print("set", flush=True)
t0 = time.time(); sc = ctx.forward_scores(); print("forward %.3f s, emis %.3f ms score %.3f ms" % (time.time() - t0, ctx.kernel_ms("emis"), ctx.kernel_ms("score")), flush=True)
lab, _ = ctx.rank(sc); print("top1", float(np.mean(lab == labels)), flush=True)
t0 = time.time(); sc = ctx.viterbi_scores(); print("viterbi scores %.3f s" % (time.time() - t0), flush=True)
ctx.close()
