"""Where the one-launch peer all-reduce (k_peer_allreduce1) spends its time: phase stamps of every block (option peer_dbg),
with the ranks released together as in bench.py's allreduce_alone_ms.  torchrun --nproc-per-node N scripts/peer_stamps.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_recognition_hmm_continuous_b200 import api, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
V, N, M, D = 10, 5, 16, 39
cen, s = synth.make_centres(V, N, M, D, seed=7)
labels = np.repeat(np.arange(V), 10)
x, off = synth.make_utterances(cen, s, labels, seed=100 + rank)
ms = api.ModelSet.from_dict(synth.make_models(cen, s))
c = api.Context(local)
c.set_features(x, off)
c.set_models(ms)
c.estep(labels, download=False, want_logp=False)
hs = [None] * world
dist.all_gather_object(hs, c.peer_export(world))
c.peer_import(rank, world, hs)
c.set_option("peer_dbg", 1)
if len(sys.argv) > 1:
    c.set_option("peer_ll", int(sys.argv[1]))  # 1 = k_peer_allreduce_ll (default), 0 = k_peer_allreduce1
st = torch.cuda.ExternalStream(c.stream(), device=local)
al = torch.zeros(1, device=dev)
lib = c.lib
lib.hmmcu_debug_peer_read.argtypes = [C.c_void_p, C.c_void_p]
rows, evs = [], []
for it in range(23):
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        dist.all_reduce(al)
        torch.cuda._sleep(200000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    c.peer_allreduce()
    e1.record(st)
    torch.cuda.synchronize()
    buf = np.zeros((128, 8), dtype=np.int64)
    assert lib.hmmcu_debug_peer_read(c.h, buf.ctypes.data_as(C.c_void_p)) == 0
    if it >= 3:
        b = buf[buf[:, 0] > 0]
        t0 = b[:, 0].min()
        rows.append([len(b), (b[:, 0].max() - t0), (b[:, 1] - b[:, 0]).mean(), (b[:, 2] - b[:, 1]).mean(), (b[:, 3] - b[:, 2]).mean(),
                     (b[:, 3] - b[:, 2]).max(), (b[:, 4] - b[:, 3]).mean(), b[:, 4].max() - t0, e0.elapsed_time(e1) * 1e6])
assert not c.peer_error()
r = np.array(rows, dtype=np.float64).mean(axis=0)
out = ("rank %d: %d blocks | last block starts +%.0f ns | stores issued %.0f | fence %.0f | wait for peers' flags mean %.0f max %.0f | "
       "sum %.0f | first start -> last end %.0f ns | event-timed %.0f ns" % ((rank,) + tuple(r)))
allo = [None] * world
dist.all_gather_object(allo, out)
if rank == 0:
    print("\n".join(allo))
c.close()
dist.destroy_process_group()
