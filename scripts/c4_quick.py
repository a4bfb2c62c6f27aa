"""C4-shaped decode slice (1,000 words, N=5, M=3, 400 utterances): emission / scorer kernel times."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 1000, 5, 3, 400
cen, s = synth.make_centres(V, N, M, 39, seed=77)
labels = (np.arange(U) % V).astype(np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=78)
ctx = api.Context(0, timing=True)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.set_features(x, off)
ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
for leg, fn, kn in (("forward", ctx.forward_scores, "score"), ("viterbi", ctx.viterbi_scores, "viterbi")):
    fn()
    em, sc = [], []
    for _ in range(5):
        sco = fn()
        em.append(ctx.kernel_ms("emis_total")); sc.append(ctx.kernel_ms(kn + "_total"))
    F = int(off[-1])
    flops = 2.0 * 79 * V * N * M * F
    print("%s: emis %.3f ms (%.1f TF/s algorithmic, x3 issued %.1f)  score %.3f ms (%.0f GB/s)  frames %d" % (leg, np.median(em), flops / np.median(em) / 1e9, 3 * flops / np.median(em) / 1e9, np.median(sc), (4.0 * N + (N if leg == "viterbi" else 0)) * V * F / np.median(sc) / 1e6, F))
lab, _ = ctx.rank(sco)
print("top1", float(np.mean(lab == labels)), "dec_grid", ctx.kernel_ms("dec_grid"))
