"""k_accum_h against k_accum_ws on the same E-step (statistics and time): python scripts/dbg_acch.py [U]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M = 10, 5, 16
U = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels_all = (np.arange(U) % V).astype(np.int32)
bench.GEN_BLOCK = U
x, off, lab = bench.gen_corpus_device(torch, dev, cen, s, labels_all, 0, U, seed=1235)
ctx = api.Context(0, timing=True)
ctx.set_features_device(x.data_ptr(), off, 39)
ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
res = {}
for h in (0, 1, 0, 1):
    ctx.set_option("h_acc", h)
    ms = []
    for _ in range(4):
        st, _ = ctx.estep(lab, download=True, want_logp=False)
        ctx.synchronize()
        ms.append(ctx.kernel_ms("accum"))
    res[h] = np.array(st)
    print("h_acc=%d accum %.3f ms (emis %.3f fwdbwd %.3f pack_x16 %.3f)" % (h, float(np.median(ms)), ctx.kernel_ms("emis"), ctx.kernel_ms("fwdbwd"), ctx.kernel_ms("pack_x16")), flush=True)
D = 39
worst = {}
for v in range(V):
    a = api.split_stats(res[0][v], N, M, D)
    b = api.split_stats(res[1][v], N, M, D)
    S0 = np.maximum(a["S0"], 1e-300)[..., None]
    sd = np.sqrt(np.maximum(a["S2c"], 0) / S0)
    for name in ("S0", "S1", "S2c", "num_trans", "den_mix"):
        if name == "S0":
            e = np.abs(b[name] - a[name]).max() / np.abs(a[name]).max()
        elif name == "S1":
            e = (np.abs(b[name] / S0 - a[name] / S0) / np.maximum(np.abs(a[name] / S0), sd)).max()
        elif name == "S2c":
            e = (np.abs(b[name] - a[name]) / np.maximum(a[name], 1e-300)).max()
        else:
            e = np.abs(b[name] - a[name]).max() / max(np.abs(a[name]).max(), 1e-300)
        worst[name] = max(worst.get(name, 0.0), float(e))
print("max relative difference h_acc=1 vs 0:", worst)
print("nan in h stats:", int(np.isnan(res[1]).sum()), " S0 sums:", float(api.split_stats(res[0][0], N, M, D)["S0"].sum()), float(api.split_stats(res[1][0], N, M, D)["S0"].sum()))
