import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
from speech_recognition_hmm_continuous_b200 import api, synth
V, N, M, U = 10, 5, 16, 1000
cen, s = synth.make_centres(V, N, M, 39, seed=1234)
labels = np.arange(U) % V
x, off = synth.make_utterances(cen, s, labels, seed=1234)
c = api.Context(0)
c.set_features(x, off)
c.enable_timing(True)
ms = c.init_models(labels, V, N, M)
c.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    c.init_models(labels, V, N, M)
print("init ms", (time.perf_counter() - t0) / 3 * 1e3, "kernel", c.kernel_ms("init"))
us0 = np.nonzero(labels == 3)[0]
x0 = np.concatenate([x[off[u]:off[u + 1]] for u in us0])
off0 = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us0])]).astype(np.int64)
h = api.init_model(N, M, x0, off0)
for k in ("A", "c", "mu", "iv", "det"):
    assert np.array_equal(getattr(ms, k)[3], getattr(h, k)[0]), k
print("bit-identical to the host builder")
