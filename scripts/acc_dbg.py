"""E-step only (no M-step feedback) at the C3-shard size with the accumulate kernel's experiment switches: python scripts/acc_dbg.py"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 10, 5, 16, 12500
dev = torch.device("cuda", 0)
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels_all = (np.arange(U) % V).astype(np.int32)
bench.GEN_BLOCK = U
x, off, lab = bench.gen_corpus_device(torch, dev, cen, s, labels_all, 0, U, seed=1235)
ctx = api.Context(0, timing=True)
ctx.set_features_device(x.data_ptr(), off, 39)
ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
for dbg in (0, 1, 2, 3, 0):
    ctx.set_option("acc_dbg", dbg)
    ms = []
    for _ in range(4):
        ctx.estep(lab, download=False, want_logp=False)
        ctx.synchronize()
        ms.append(ctx.kernel_ms("accum"))
    print("acc_dbg=%d accum %.3f ms (emis %.3f fwdbwd %.3f)" % (dbg, float(np.median(ms)), ctx.kernel_ms("emis"), ctx.kernel_ms("fwdbwd")))
