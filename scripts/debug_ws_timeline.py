"""Timeline of CTA 0 of the warp-specialised emission kernel (clock64 stamps per unit and role)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "train"
V, N, M, U = (1000, 5, 3, 50) if mode == "dec4" else (100, 3, 128, 50) if mode == "dec5" else (10, 5, 16, 1000)
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels = (np.arange(U) % V).astype(np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=1234)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_features(x, off); ctx.set_models(ms); ctx.em_reset()
for _ in range(3 if mode in ("train", "acc") else 0):
    ctx.estep(labels, download=False, want_logp=False)
ctx.set_option("debug_acc", 4 if mode == "acc" else 2)
if mode in ("train", "acc"):
    ctx.estep(labels, download=False, want_logp=False)
else:
    ctx.forward_scores()
ctx.synchronize()
buf = np.zeros(3 * 16384, dtype=np.float32)
ctx.lib.hmmcu_debug_acc_read(ctx.h, buf.ctypes.data_as(C.c_void_p))
W = 16 if mode == "acc" else 8
t = buf.view(np.int64)[:64 * W].reshape(64, W)[:, :15 if mode == "acc" else 8]
t0 = t[t > 0].min()
names = ["ld:top", "ld:empty", "ld:arrive", "mma:full", "mma:dempty", "mma:issued", "epi:dfull", "epi:arrive"]
if mode == "acc":
    names = ["ld:top", "ld:free", "ld:arrive", "mma:xfull", "mma:g1", "mma:g2", "epi:d1full", "epi:wfull", "g2:start", "g2:issued", "x:stored", "x:cfs", "xt:free", "xt:arrive", "xt:top"]
print("unit " + " ".join("%10s" % n for n in names))
for i in range(64):
    if t[i].max() == 0: break
    print("%4d " % i + " ".join("%10d" % (v - t0 if v > 0 else -1) for v in t[i]))

if mode == "acc":
    g = buf.view(np.int64)
    print("kernel entry %d, after init %d, role ends %s, after final sync %d" % (g[1024] - t0, g[1025] - t0, [int(v - t0) for v in g[1026:1046]], g[1050] - t0))
