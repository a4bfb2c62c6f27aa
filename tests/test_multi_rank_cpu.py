"""The N > 1 path on CPU: two gloo ranks shard the utterances, each forms the sufficient statistics of its shard
(the oracle stands in for the CUDA E-step, which needs a GPU), the statistics vector -- in the library's layout,
the payload of the one all-reduce per EM iteration -- is summed across the ranks, and every rank runs
the same host M-step (hmmh_mstep).  Both ranks must end with bit-identical models, equal (up to summation order)
to the single-process result."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle as o
from speech_recognition_hmm_continuous_b200 import api, synth

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

V, N, M, U, D = 2, 4, 2, 10, 39


def _data():
    cen, s = synth.make_centres(V, N, M, D, seed=5150)
    labels = np.arange(U) % V
    x, off = synth.make_utterances(cen, s, labels, seed=5151, tmin=30, tmax=60)
    rng = np.random.default_rng(9)
    ms = api.ModelSet.from_dict(synth.make_models(cen + 0.2 * s * rng.standard_normal(cen.shape), s))
    return x, off, labels, ms


def _stats_of(ms, x, off, labels, u0, u1):
    """[V][stats_size] of utterances [u0, u1), in the layout of hmmcu_stats_size."""
    ss = api.stats_size(N, M, D)
    out = np.zeros((V, ss))
    for v in range(V):
        us = [u for u in range(u0, u1) if labels[u] == v]
        if not us:
            continue
        xv = np.concatenate([x[off[u]:off[u + 1]] for u in us])
        offv = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])]).astype(np.int64)
        st, _ = o.estep(o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v]), xv, offv)
        out[v] = np.concatenate([st.num_trans.ravel(), st.den_trans, st.den_mix, st.S0.ravel(), st.S1.ravel(), st.S2c.ravel(),
                                 [st.sum_logp, st.n_utt]])
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, off, labels, ms = _data()
    u0, u1 = api.shard_utterances(off, rank, world)
    t = torch.from_numpy(_stats_of(ms, x, off, labels, u0, u1))
    dist.all_reduce(t)  # the one collective of an EM iteration
    stats = t.numpy()
    api.mstep(ms, stats)
    q.put((rank, u0, u1, stats.copy(), ms.A.copy(), ms.c.copy(), ms.mu.copy(), ms.iv.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_utterances_is_a_balanced_partition():
    _, off, _, _ = _data()
    for world in (1, 2, 3, 8):
        cuts = [api.shard_utterances(off, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == U
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        frames = [off[b] - off[a] for a, b in cuts]
        if world <= 3:
            assert max(frames) - min(frames) <= 2 * 60  # within two utterances of each other


def test_two_gloo_ranks_reproduce_the_single_process_em_iteration():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, st0, A0, c0, mu0, iv0), (r1, a1, b1, st1, A1, c1, mu1, iv1) = res
    assert (a0, b1) == (0, U) and b0 == a1 and 0 < b0 < U
    # every rank holds the same sum and therefore the same models, bit for bit
    assert np.array_equal(st0, st1)
    for x0, x1 in ((A0, A1), (c0, c1), (mu0, mu1), (iv0, iv1)):
        assert np.array_equal(x0, x1)
    # and it is the single-process iteration up to the order of the floating-point sums
    x, off, labels, ms = _data()
    full = _stats_of(ms, x, off, labels, 0, U)
    assert np.allclose(st0, full, rtol=1e-11, atol=1e-9)
    api.mstep(ms, full)
    for got, want in ((A0, ms.A), (c0, ms.c), (mu0, ms.mu), (iv0, ms.iv)):
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)


# ---- two feature streams: every stream's statistics vector is all-reduced, every rank runs both M-steps ----
STREAMS = ((2, 39), (3, 13))  # (M, D) of the two streams


def _data_streams():
    labels = np.arange(U) % V
    sets = []
    for p, (Mp, Dp) in enumerate(STREAMS):
        cen, s = synth.make_centres(V, N, Mp, Dp, seed=6100 + p)
        x, off = synth.make_utterances(cen, s, labels, seed=6151, tmin=30, tmax=60)   # the same seed: the same lengths
        rng = np.random.default_rng(19 + p)
        sets.append((api.ModelSet.from_dict(synth.make_models(cen + 0.2 * s * rng.standard_normal(cen.shape), s)), x, off))
    assert np.array_equal(sets[0][2], sets[1][2])
    return sets, labels


def _stream_stats(sets, labels, u0, u1):
    """[[V][stats_size_p] per stream] of utterances [u0, u1): what hmmcu_estep leaves in the linked contexts."""
    off = sets[0][2]
    out = [np.zeros((V, api.stats_size(N, Mp, Dp))) for Mp, Dp in STREAMS]
    for v in range(V):
        us = [u for u in range(u0, u1) if labels[u] == v]
        if not us:
            continue
        offv = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us])]).astype(np.int64)
        xs = [np.concatenate([x[off[u]:off[u + 1]] for u in us]) for _, x, _ in sets]
        models = [o.Model(ms.A[v], ms.c[v], ms.mu[v], ms.iv[v], ms.det[v]) for ms, _, _ in sets]
        sts, _ = o.estep_streams(models, xs, offv)
        for p, st in enumerate(sts):
            out[p][v] = np.concatenate([st.num_trans.ravel(), st.den_trans, st.den_mix, st.S0.ravel(), st.S1.ravel(), st.S2c.ravel(),
                                        [st.sum_logp, st.n_utt]])
    return out


def _worker_streams(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sets, labels = _data_streams()
    u0, u1 = api.shard_utterances(sets[0][2], rank, world)
    out = []
    for p, st in enumerate(_stream_stats(sets, labels, u0, u1)):
        t = torch.from_numpy(st)
        dist.all_reduce(t)  # hmmh_train_streams: the all-reduce hook runs once per stream and iteration
        api.mstep(sets[p][0], t.numpy())
        out.append((t.numpy().copy(), sets[p][0].A.copy(), sets[p][0].mu.copy(), sets[p][0].iv.copy()))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_two_streams():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_streams, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sets, labels = _data_streams()
    full = _stream_stats(sets, labels, 0, U)
    head = N * N + 2 * N
    for p in range(2):
        (st0, A0, mu0, iv0), (st1, A1, mu1, iv1) = res[0][1][p], res[1][1][p]
        assert np.array_equal(st0, st1) and np.array_equal(mu0, mu1) and np.array_equal(A0, A1)
        assert np.allclose(st0, full[p], rtol=1e-11, atol=1e-9)
        api.mstep(sets[p][0], full[p])
        assert np.allclose(mu0, sets[p][0].mu, rtol=1e-9, atol=1e-12) and np.allclose(iv0, sets[p][0].iv, rtol=1e-9)
    # the transition statistics (and therefore A) are the same in both streams' vectors
    assert np.array_equal(res[0][1][0][0][:, :head], res[0][1][1][0][:, :head]) and np.array_equal(res[0][1][0][1], res[0][1][1][1])
