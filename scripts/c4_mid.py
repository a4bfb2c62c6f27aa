"""A larger C4-shaped decode leg (1,000 words, N=5, M=3, 4,000 utterances, features generated on the device)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 1000, 5, 3, int(os.environ.get("U", "4000"))
dev = torch.device("cuda", 0)
cen, s = synth.make_centres(V, N, M, 39, seed=4004)
labels_all = (np.arange(U) % V).astype(np.int32)
bench.GEN_BLOCK = U
x, off, lab = bench.gen_corpus_device(torch, dev, cen, s, labels_all, 0, U, seed=4005)
TIMING = os.environ.get("TIMING", "1") == "1"
ctx = api.Context(0, timing=TIMING)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.set_features_device(x.data_ptr(), off, 39)
ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
F = int(off[-1])
flops = 2.0 * 79 * V * N * M * F
ctx.forward_scores()
for rep in range(3):
    t0 = time.perf_counter(); sco = ctx.forward_scores(); wall = time.perf_counter() - t0
    if not TIMING:
        print("forward: wall %.1f ms frames %d (%.1f M frames/s)" % (wall * 1e3, F, F / wall / 1e6))
        continue
    em, sc = ctx.kernel_ms("emis_total"), ctx.kernel_ms("score_total")
    print("forward: emis %.2f ms (%.1f TF/s algorithmic, x3 = %.1f) score %.2f ms wall %.1f ms frames %d" % (em, flops / em / 1e9, 3 * flops / em / 1e9, sc, wall * 1e3, F))
labg, _ = ctx.rank(sco)
print("top1", float(np.mean(labg == lab)))
