"""C4-shape decode in a loop (clock / power sampling beside it): python scripts/c4_loop.py [key=value ...]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 1000, 5, 3, 4000
dev = torch.device("cuda", 0)
cen, s = synth.make_centres(V, N, M, 39, seed=4004)
labels_all = (np.arange(U) % V).astype(np.int32)
bench.GEN_BLOCK = U
x, off, lab = bench.gen_corpus_device(torch, dev, cen, s, labels_all, 0, U, seed=4005)
ctx = api.Context(0, timing=True)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.set_features_device(x.data_ptr(), off, 39)
ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
t0 = time.time()
ems = []
while time.time() - t0 < 6.0:
    ctx.forward_scores()
    ems.append(ctx.kernel_ms("emis_total"))
print("emis ms: first %.2f median %.2f last %.2f (%d calls)" % (ems[0], float(np.median(ems)), ems[-1], len(ems)))
