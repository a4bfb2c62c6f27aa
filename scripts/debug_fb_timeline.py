"""Phase time stamps (globaltimer, ns) of the resident forward-backward kernel k_fb_res, per team:
   python scripts/debug_fb_timeline.py [c2|c3]      -> durations of fetch / staging 1 / staging 2 / chains / gamma phase
   python scripts/debug_fb_timeline.py n2            -> the N = 2 case of the parity tests, resident vs windowed kernel"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "c2"
if mode == "n2":
    from oracle import oracle as o
    for N in (2, 4):
        cen, s = synth.make_centres(2, N, 2, 39, seed=880 + N)
        labels = np.arange(7) % 2
        x, off = synth.make_utterances(cen, s, labels, seed=881 + N, tmin=2, tmax=150)
        ms = api.ModelSet.from_dict(synth.make_models(cen, s))
        for res in (1, 0):
            ctx = api.Context(0)
            ctx.set_option("res_fb", res)
            ctx.set_features(x, off); ctx.set_models(ms)
            st, lp = ctx.estep(labels)
            lb, _ = ctx.emissions(0, 0)
            mo = o.Model(ms.A[0], ms.c[0], ms.mu[0], ms.iv[0], ms.det[0])
            b, _ = o.emissions(mo, x[off[0]:off[1]], want_post=False)
            print("N", N, "res", res, "T", np.diff(off), "lp", lp, "emis maxdiff", np.abs(lb - np.log(b)).max())
            ctx.close()
    sys.exit(0)
V, N, M, U = (10, 5, 16, 12500) if mode == "c3" else (10, 5, 16, 1000)
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels = (np.arange(U) % V).astype(np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=1234)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_features(x, off); ctx.set_models(ms); ctx.em_reset()
for _ in range(3):
    ctx.estep(labels, download=False, want_logp=False)
ctx.set_option("debug_acc", 8)
ctx.enable_timing(True)
ctx.estep(labels, download=False, want_logp=False)
ctx.synchronize()
print("fwdbwd ms", ctx.kernel_ms("fwdbwd"))
buf = np.zeros(3 * 16384, dtype=np.float32)
ctx.lib.hmmcu_debug_acc_read(ctx.h, buf.ctypes.data_as(C.c_void_p))
t = buf.view(np.int64)[:148 * 4 * 32].reshape(148 * 4, 32)
live = t[:, 0] > 0
t0 = t[live, 0].min()
print("teams with work:", int(live.sum()))
names = ["fetch+meta+stage1", "stage2", "chains", "gamma"]
# stamps per batch: top, after stage 1, after stage 2, after chains; the next batch's top closes the gamma phase
r0 = t[live][0]
print("SM clock during the kernel: %.0f MHz (clock64 / globaltimer over the first team's stamps)" % ((r0[16 + 4] - r0[16]) / max(r0[4] - r0[0], 1) * 1e3))
d = []
for row in t[live]:
    k = 0
    while k + 4 < 16 and row[k + 4] > 0:
        d.append([row[k + 1] - row[k], row[k + 2] - row[k + 1], row[k + 3] - row[k + 2], row[k + 4] - row[k + 3]])
        k += 4
d = np.array(d, dtype=np.float64)
print("batches timed:", len(d))
for i, n in enumerate(names):
    print("%-20s mean %8.0f ns   min %8.0f   max %8.0f" % (n, d[:, i].mean(), d[:, i].min(), d[:, i].max()))
print("first stamps (ns after the first): ", [int(v - t0) for v in t[live][0][:9]])
print("last stamp overall:", int(t[live].max() - t0))
