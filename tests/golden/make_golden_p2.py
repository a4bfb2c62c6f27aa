"""Generates tests/golden/synth_p2.npz by RUNNING THE REFERENCE ITSELF with two feature streams
(param_number = 2; oracle/_ref/*_p2, built from /root/reference by oracle/build_ref.sh with
MAX_PARAMETERS_NUMBER 2).  Run in the build container only:

    python tests/golden/make_golden_p2.py

Case: V = 2 words, N = 4 states; stream 0 has D = 6, M = 2; stream 1 has D = 4, M = 3; 8 training and 3 test
utterances per word, 40..60 frames.  Stored: the features of both streams, the reference trainer's models
(both streams), its iterations and mean log-probabilities, and the recogniser's sorted candidate lists with
their scores as printed ("%f").
"""
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as r  # noqa: E402
from speech_recognition_hmm_continuous_b200 import synth  # noqa: E402

G = os.path.dirname(os.path.abspath(__file__))
V, N = 2, 4
STREAMS = ((2, 6), (3, 4))  # (M, D)


def main():
    if not r.available("p2"):
        sys.exit("oracle/_ref/*_p2 missing: run oracle/build_ref.sh first")
    tmp = tempfile.mkdtemp()
    train_labels = np.repeat(np.arange(V), 8)
    test_labels = np.repeat(np.arange(V), 3)
    data = {}
    for p, (M, D) in enumerate(STREAMS):
        cen, s = synth.make_centres(V, N, M, D, seed=500 + p)
        # the same seed draws the same utterance lengths in every stream
        data["x%d" % p], off = synth.make_utterances(cen, s, train_labels, seed=77, tmin=40, tmax=60)
        data["xt%d" % p], offt = synth.make_utterances(cen, s, test_labels, seed=78, tmin=40, tmax=60)
    exe = os.path.join(r.REF_DIR, "hmm_fs_p2")
    models, means, iters = [], [], []
    import subprocess
    for v in range(V):
        lists = []
        for p in range(2):
            files = []
            for u in np.nonzero(train_labels == v)[0]:
                f = os.path.join(tmp, "tr_s%d_%d.bin" % (p, u))
                r.write_features(f, data["x%d" % p][off[u]:off[u + 1]])
                files.append(f)
            lists.append(os.path.join(tmp, "list_s%d_w%d.txt" % (p, v)))
            open(lists[-1], "w").write("\n".join(files) + "\n")
        hmm = os.path.join(tmp, "w%d.hmm" % v)
        subprocess.run([exe, "word%d" % v, str(N), "2", str(STREAMS[0][0]), str(STREAMS[1][0]), lists[0], lists[1], hmm],
                       stdout=subprocess.DEVNULL, check=True)
        mean, its = r.parse_train_report(hmm[:-4] + ".txt")
        models.append(r.read_model_streams(hmm)); means.append(mean); iters.append(its)
    feats = []
    for p in range(2):
        files = []
        for u in range(len(test_labels)):
            f = os.path.join(tmp, "te_s%d_%d.bin" % (p, u))
            r.write_features(f, data["xt%d" % p][offt[u]:offt[u + 1]])
            files.append(f)
        feats.append(os.path.join(tmp, "feat_s%d.txt" % p))
        open(feats[-1], "w").write("\n".join(files) + "\n")
    open(os.path.join(tmp, "models.txt"), "w").write("\n".join(os.path.join(tmp, "w%d.hmm" % v) for v in range(V)) + "\n")
    open(os.path.join(tmp, "words.txt"), "w").write("\n".join("word%d" % v for v in test_labels) + "\n")
    out = subprocess.run([os.path.join(r.REF_DIR, "rec_fs_p2"), "1", os.path.join(tmp, "models.txt"), "1", feats[0], feats[1],
                          os.path.join(tmp, "words.txt"), os.path.join(tmp, "res.txt")], stdout=subprocess.PIPE, check=True).stdout.decode(errors="replace")
    score = np.zeros((len(test_labels), V))
    for u, blk in enumerate(out.split("Spoken word: ")[1:]):
        tail = blk.split("Writing result")[1]
        for w, sc in re.findall(r"^(\S+) :  (\S+) $", tail, flags=re.M)[:V]:
            score[u, int(w[4:])] = float(sc)
    arrays = {}
    for p in range(2):
        for k in ("A", "c", "mu", "iv", "det"):
            arrays["trained_s%d_%s" % (p, k)] = np.stack([getattr(models[v][p], k) for v in range(V)])
    np.savez_compressed(os.path.join(G, "synth_p2.npz"), V=V, N=N, M=np.array([s[0] for s in STREAMS]), D=np.array([s[1] for s in STREAMS]),
                        train_labels=train_labels, test_labels=test_labels, off=off, offt=offt, mean_logp=np.array(means),
                        iterations=np.array(iters), score=score, result_file=np.array(open(os.path.join(tmp, "res.txt")).read()), **data, **arrays)
    print("synth_p2: iterations", iters, "mean logP", means, "scores\n", score)


if __name__ == "__main__":
    main()
