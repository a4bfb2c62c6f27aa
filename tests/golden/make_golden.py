"""Generates tests/golden/* by RUNNING THE REFERENCE ITSELF (oracle/_ref, built from
/root/reference by oracle/build_ref.sh).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs
  perfil/*.perfil      the reference's 13 shipped feature files (data, D=9, T=103..213)
  kat_diag.json        reference trainer (T-FS) on each of them: N=6, M=1 -> mean logP, iterations;
                       reference recogniser (R-FS) on the 13 resulting models: sorted candidates,
                       exactly as printed (this is the degenerate NaN / -inf regime)
  kat_models.npz       the 13 trained diagonal models (float64) written by the reference trainer
  synth_c1.npz         finite-regime case (synthetic, D=39, N=5, M=3, V=6): reference-trained
                       models, per-word mean logP / iterations, function-level E-step statistics
                       of iteration 1, full-precision forward-score matrix, ranking, labels
"""
import glob
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as r  # noqa: E402
from speech_recognition_hmm_continuous_b200 import synth  # noqa: E402

G = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def model_arrays(models, prefix=""):
    out = {}
    for k in ("A", "c", "mu", "iv", "det"):
        out[prefix + k] = np.stack([getattr(m, k) for m in models])
    return out


def shipped_kats():
    tmp = tempfile.mkdtemp()
    files = sorted(glob.glob(os.path.join(REF, "train/test/perfil_data/*.perfil")))
    kat = {"train": {}, "recognition": []}
    models, words = [], []
    os.makedirs(os.path.join(tmp, "models"))
    for p in files:
        shutil.copy(p, os.path.join(G, "perfil", os.path.basename(p)))
        word = os.path.basename(p)[len("mean_"):-len(".perfil")]
        lst = os.path.join(tmp, "l.txt")
        open(lst, "w").write(p + "\n")
        hmm = os.path.join(tmp, "models", "mean_%s.hmm" % word)
        r.run_train_cli("stock", word, 6, 1, lst, hmm)
        mean, its = r.parse_train_report(hmm[:-4] + ".txt")
        kat["train"][word] = {"mean_logp": mean, "iterations": its, "file": os.path.basename(p)}
        models.append(r.read_model(hmm))
        words.append(word)
    # recogniser on the 13 diagonal models, same order as the shipped lists
    open(os.path.join(tmp, "models.txt"), "w").write("\n".join(os.path.join(tmp, "models", "mean_%s.hmm" % w) for w in words) + "\n")
    open(os.path.join(tmp, "feat.txt"), "w").write("\n".join(files) + "\n")
    open(os.path.join(tmp, "words.txt"), "w").write("\n".join(words) + "\n")
    out = r.run_test_cli("stock", os.path.join(tmp, "models.txt"), os.path.join(tmp, "feat.txt"),
                         os.path.join(tmp, "words.txt"), os.path.join(tmp, "res.txt"), capture=True)
    blocks = out.split("Spoken word: ")[1:]
    for blk in blocks:
        spoken = blk.split()[0]
        tail = blk.split("Writing result")[1]
        cands = re.findall(r"^(\S+) :  (\S+) $", tail, flags=re.M)
        kat["recognition"].append({"spoken": spoken, "sorted": cands})
    kat["result_file"] = open(os.path.join(tmp, "res.txt")).read()
    json.dump(kat, open(os.path.join(G, "kat_diag.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(G, "kat_models.npz"), words=np.array(words), **model_arrays(models))
    print("shipped KATs:", len(files), "files")


def synth_c1():
    V, N, M, D = 6, 5, 3, 39
    tmp = tempfile.mkdtemp()
    cen, s = synth.make_centres(V, N, M, D, seed=1234)
    train_labels = np.repeat(np.arange(V), 4)
    test_labels = np.repeat(np.arange(V), 2)
    x, off = synth.make_utterances(cen, s, train_labels, seed=1234, tmin=70, tmax=110)
    xt, offt = synth.make_utterances(cen, s, test_labels, seed=4321, tmin=70, tmax=110)
    rt = r.RefTrain("d39m16")
    models, init_models, means, iters = [], [], [], []
    stats = {k: [] for k in ("num_trans", "den_trans", "den_mix", "S0", "S1", "S2c", "sum_logp")}
    for v in range(V):
        files = []
        for u in np.nonzero(train_labels == v)[0]:
            f = os.path.join(tmp, "tr_%d.bin" % u)
            r.write_features(f, x[off[u]:off[u + 1]])
            files.append(f)
        lst = os.path.join(tmp, "list_%d.txt" % v)
        open(lst, "w").write("\n".join(files) + "\n")
        hmm = os.path.join(tmp, "w%d.hmm" % v)
        r.run_train_cli("d39m16", "word%d" % v, N, M, lst, hmm)
        mean, its = r.parse_train_report(hmm[:-4] + ".txt")
        models.append(r.read_model(hmm)); means.append(mean); iters.append(its)
        mi = rt.init_model(N, M, lst, "word%d" % v)
        init_models.append(mi)
        acc = rt.new_acc(); tot = 0.0
        for u in np.nonzero(train_labels == v)[0]:
            tot += rt.utterance(mi, x[off[u]:off[u + 1]], acc)["logp"]
        st = rt.acc_to_stats(acc, N, M, D)
        for k in ("num_trans", "den_trans", "den_mix", "S0", "S1", "S2c"):
            stats[k].append(getattr(st, k))
        stats["sum_logp"].append(tot)
    rte = r.RefTest("d39m16")
    U = len(test_labels)
    score = np.zeros((U, V))
    for u in range(U):
        for v in range(V):
            score[u, v] = rte.forward_score(models[v], xt[offt[u]:offt[u + 1]])
    order = np.stack([rte.rank(score[u]) for u in range(U)])
    np.savez_compressed(
        os.path.join(G, "synth_c1.npz"), V=V, N=N, M=M, D=D, train_labels=train_labels, test_labels=test_labels,
        mean_logp=np.array(means), iterations=np.array(iters), score=score, order=order,
        **model_arrays(models, "trained_"), **model_arrays(init_models, "init_"),
        **{"stat_" + k: np.stack(v) for k, v in stats.items()})
    print("synth_c1: labels", order[:, 0], "truth", test_labels, "iters", iters)


if __name__ == "__main__":
    if not r.available("stock"):
        sys.exit("oracle/_ref missing: run oracle/build_ref.sh first")
    shipped_kats()
    synth_c1()
