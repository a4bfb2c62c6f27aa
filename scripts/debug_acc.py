"""Diagnostic for the tensor-core accumulate kernel: dumps the first unit's intermediates."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth
from oracle import oracle as o
V, N, M, U = 1, 5, 3, 2
cen, s = synth.make_centres(V, N, M, 39, seed=1)
labels = np.zeros(U, dtype=np.int32)
x, off = synth.make_utterances(cen, s, labels, seed=2, tmin=70, tmax=90)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
ctx = api.Context(0)
ctx.set_option("tc_emis", 2)
ctx.set_option("debug_acc", 1)
ctx.set_features(x, off)
ctx.set_models(ms)
stats, lpu = ctx.estep(labels)
buf = np.zeros(3 * 16384, dtype=np.float32)
rc = ctx.lib.hmmcu_debug_acc_read(ctx.h, buf.ctypes.data_as(C.c_void_p))
print("rc", rc)
L2 = buf[:16384].reshape(128, 128); W = buf[16384:32768].reshape(128, 128); D2 = buf[32768:].reshape(128, 128)
G = N * M
# expected log2(c_g N_g(x_f)) for the first 128 frames
mu = ms.mu[0].reshape(G, 39); iv = ms.iv[0].reshape(G, 39); det = ms.det[0].reshape(G); c = ms.c[0].reshape(G)
xf = x[:128]
q = ((xf[None, :, :] - mu[:, None, :]) ** 2 * iv[:, None, :]).sum(-1)
ln = np.log(c)[:, None] - 0.5 * q - 0.5 * (39 * np.log(2 * np.pi) + np.log(det))[:, None]
want = ln / np.log(2)
print("L2 got[:3,:4]\n", L2[:3, :4], "\nwant\n", want[:3, :4])
print("L2 max abs err over G x 128:", np.abs(L2[:G] - want).max())
print("W colsum (first 8 frames):", W[:G].sum(0)[:8])
print("W max", W.max(), "nonzero", (W != 0).sum())
ns = min(len(x), 4096); stride = len(x) // ns
ctr = x[np.arange(ns) * stride].mean(0)
xc = (xf - ctr).astype(np.float32)
Xaug = np.concatenate([xc, np.ones((128, 1), np.float32), xc ** 2, np.ones((128, 1), np.float32)], 1)
want2 = W[:G].astype(np.float64) @ Xaug.astype(np.float64)
print("D2 got[:2,:6]\n", D2[:2, :6], "\nwant\n", want2[:2, :6])
print("D2 S0 col got", D2[:G, 39][:6], "want", want2[:, 39][:6])
print("D2 max abs err", np.abs(D2[:G, :80] - want2).max(), "scale", np.abs(want2).max())
sp = api.split_stats(stats[0], N, M, 39)
print("S0", sp["S0"])
