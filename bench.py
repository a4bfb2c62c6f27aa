#!/usr/bin/env python
"""bench.py -- frames/s of one Baum-Welch EM iteration (and of decode) on B200, next to the
reference's own CPU implementation.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c5|c4slice] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU; utterances sharded, one all-reduce of the
                                             sufficient statistics per EM iteration)

A "step" is one EM iteration over the rank's utterances -- the loop body of the reference trainer's
main() (T-FS:238-358) through the C ABI: hmmcu_estep (emissions, forward-backward, accumulators),
the all-reduce of the statistics, hmmcu_mstep (device M-step + the stopping rule; its 3V+1 doubles
are read back every step).  The stopping threshold is set to -1 so that every word is re-estimated
in every step and the work per step stays the full workload.  `value` times it with the features
resident in HBM; `e2e` additionally re-uploads the double-precision features from pinned host
memory every step (hmmcu_set_features), i.e. what the drop-in trainer pays on its first iteration.

Default workload = BASELINE.json configs[1] ("c2"): 5-state left-to-right HMMs, 16 mixtures/state,
39-dim frames, 10 words, 1,000 utterances of ~300 frames per GPU.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION (some images export that); stdout carries the one JSON line
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: V words, N states, M mixtures, U utterances per GPU, description
    "c2": dict(V=10, N=5, M=16, U=1000, desc="BASELINE configs[1]: N=5 M=16 D=39, 10 words, 1000 utterances (~300 frames) per GPU"),
    "c1": dict(V=10, N=5, M=3, U=220, desc="BASELINE configs[0]: N=5 M=3 D=39, 10 words, 22 utterances per word"),
    "c3": dict(V=10, N=5, M=16, U=12500, desc="BASELINE configs[2] shard: N=5 M=16 D=39, 10 words, 12,500 utterances (100k over 8 GPUs) per GPU"),
    "c5": dict(V=20, N=3, M=128, U=1000, desc="BASELINE configs[4] slice: N=3 M=128 D=39, 20 of 2000 models, 1000 utterances"),
}
D = 39
K_AUG = 2 * D + 1  # [x, x^2, 1]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml, 20 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_workload(w, rank, seed=1234):
    from speech_recognition_hmm_continuous_b200 import synth
    cen, s = synth.make_centres(w["V"], w["N"], w["M"], D, seed=seed)
    labels = (np.arange(w["U"]) % w["V"]).astype(np.int32)
    x, off = synth.make_utterances(cen, s, labels, seed=seed + 1000 * rank)
    # initial models: the generating centres perturbed, so that EM has real work to do
    rng = np.random.default_rng(seed + 7)
    mods = synth.make_models(cen + 0.3 * s * rng.standard_normal(cen.shape), s)
    return x, off, labels, mods


# ================================================================================ our arm ====
def run_ours(args):
    import torch
    import torch.distributed as dist
    from speech_recognition_hmm_continuous_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa = None
    if not args.no_affinity:
        # keep this rank (and therefore the first touch of its host buffers, pinned staging included) on the CPUs
        # of the NUMA node its GPU hangs off: with eight ranks uploading at once, remote-node pinned memory halves
        # the host-to-device rate
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = sorted(os.sched_getaffinity(0))
            numa = "%d cpus [%d..%d]" % (len(numa), numa[0], numa[-1])
        except Exception as e:  # noqa: BLE001
            numa = "unavailable: %s" % e
    w = WORKLOADS[args.workload]
    x, off, labels, mods = make_workload(w, rank)
    F = int(off[-1])
    N, M, V, U = w["N"], w["M"], w["V"], w["U"]
    G = N * M

    ctx = api.Context(local, timing=False)
    if args.upload_chunks:
        ctx.set_option("upload_chunks", args.upload_chunks)
    for kv in args.set:  # library A/B switches (hmmcu_set_option), experiments only
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    ext = torch.cuda.ExternalStream(ctx.stream(), device=local)
    ms = api.ModelSet.from_dict(mods)

    class _Alias:  # device statistics buffer as a torch tensor (no copy) for the NCCL all-reduce
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def allreduce(dev_ptr, n, stream_ptr):
        t = torch.as_tensor(_Alias(dev_ptr, n), device=torch.device("cuda", local))
        with torch.cuda.stream(ext):
            dist.all_reduce(t)

    ar = allreduce if world > 1 else None
    ar_kind = "torch.distributed (nccl)"
    if world > 1 and not args.torch_allreduce:
        try:  # the collective on the context's own stream through the NCCL C API
            ar = api.NcclAllReduce(rank, world)
            ar_kind = "ncclAllReduce on the context's stream"
        except Exception as e:  # noqa: BLE001
            print("direct NCCL unavailable (%s); using torch.distributed" % e, file=sys.stderr)
    # pinned host copy of the features (e2e leg) and a device-resident copy (value leg)
    xpin = torch.from_numpy(x).pin_memory()
    xdev = xpin.to(torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    torch.cuda.synchronize()

    def em_iteration():
        ctx.estep(labels, download=False, want_logp=False)
        if ar is not None:
            p, n = ctx.stats_device()
            ar(p, n, ctx.stream())
        return ctx.mstep(threshold=-1.0)  # reads sum_logp / n_utt / updated back: the step's result

    def step_resident():
        em_iteration()

    def step_e2e():
        ctx.set_features_ptr(xpin.data_ptr(), off, D)
        em_iteration()

    def timed(fn, steps, warmup, kernel_names=()):
        for _ in range(warmup):
            fn()
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = ctx.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kms = {k: [] for k in kernel_names}
        wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations (256 MiB > 126 MB L2), outside the event pair
            torch.cuda.synchronize()
            ev[i][0].record(ext)
            fn()
            ev[i][1].record(ext)
            ctx.synchronize()
            for k in kernel_names:
                kms[k].append(ctx.kernel_ms(k))
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        if world > 1:
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([total_ms], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count() - l0, {k: float(np.mean(v)) for k, v in kms.items()}, wall

    ctx.set_features_device(xdev.data_ptr(), off, D)
    ctx.set_models(ms)
    ctx.em_reset()
    sampler = ClockSampler(local)
    sampler.start()
    names = ("emis", "fwdbwd", "accum", "mstep")
    tot_ms, launches, _, wall = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    # per-kernel device times: a second, shorter pass with the library's event timers on (they record events
    # between the kernels, which turns the CUDA-graph replay of the iteration off -- so not the pass `value` uses)
    ctx.enable_timing(True)
    _, _, kms, _ = timed(step_resident, max(3, min(args.steps, 20)), 3, names)
    ctx.enable_timing(False)
    e2e_steps = max(3, min(args.steps, 50))
    e2e_ms, _, _, _ = timed(step_e2e, e2e_steps, max(3, min(args.warmup, 5)))

    # decode leg (forward scoring of every utterance against all V models + ranking; then Viterbi)
    ctx.set_features_device(xdev.data_ptr(), off, D)
    ctx.set_models(ms)
    ctx.enable_timing(True)  # per-kernel timers for the decode legs (no graph replay involved there)
    dec = {}
    for name, fn in (("forward", lambda: ctx.rank(ctx.forward_scores())), ("viterbi", lambda: ctx.viterbi(labels))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        ctx.synchronize()
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        ctx.synchronize()
        ms_dec = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms_dec], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kn = ("emis", "score") if name == "forward" else ("logb64", "viterbi")
        dec[name] = {"frames_per_s": F * world / (float(t.item()) * 1e-3), "ms": float(t.item()),
                     "kernel_ms": {k: ctx.kernel_ms(k) for k in kn},
                     "frame_model_pairs_per_s": (F * V * world / (float(t.item()) * 1e-3)) if name == "forward" else None}

    # initial-model builder (creating_initial_model, T-FS:732-1317): all words on the device vs the host builder
    # (single-threaded C, one word after the other) -- what the drop-in trainer pays before its first EM iteration
    init = None
    if world == 1:
        ctx.set_features_device(xdev.data_ptr(), off, D)
        ctx.init_models(labels, V, N, M)
        ctx.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ctx.init_models(labels, V, N, M)
        dev_ms = (time.perf_counter() - t0) / reps * 1e3
        us0 = np.nonzero(labels == 0)[0]
        x0 = np.concatenate([x[off[u]:off[u + 1]] for u in us0])
        off0 = np.concatenate([[0], np.cumsum([off[u + 1] - off[u] for u in us0])]).astype(np.int64)
        t0 = time.perf_counter()
        api.init_model(N, M, x0, off0)
        host_ms_word = (time.perf_counter() - t0) * 1e3
        init = {"device_ms_all_words": dev_ms, "host_ms_one_word": host_ms_word, "host_ms_all_words_extrapolated": host_ms_word * V,
                "note": "wall clock incl. the read-back of the models; results are bit-identical (tests)"}
        ctx.set_models(ms)
    # ingest: the workload's utterances as feature files -> HBM.  Serial = one read per file, concatenate,
    # hmmcu_set_features (what a straightforward host does; the reference itself issues one fread per FRAME and re-reads
    # every file twice per iteration).  Pipelined = hmmh_ingest: reader pool -> pinned staging -> async copies.
    ingest = None
    if world == 1 and not args.no_ingest:
        import shutil
        import tempfile
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        tmpd = tempfile.mkdtemp(prefix="hmmcu_ingest_", dir=base)
        try:
            paths = []
            for u in range(U):
                pth = os.path.join(tmpd, "u%06d.bin" % u)
                api.write_features(pth, x[off[u]:off[u + 1]])
                paths.append(pth)
            def serial():
                xs = [api.read_features(pth) for pth in paths]
                xo = np.concatenate(xs)
                oo = np.concatenate([[0], np.cumsum([len(a) for a in xs])]).astype(np.int64)
                ctx.set_features(xo, oo)
                ctx.synchronize()
            def piped():
                st = ctx.ingest(paths)[2]
                ctx.synchronize()
                return st
            serial(); piped()
            t0 = time.perf_counter(); serial(); ser_ms = (time.perf_counter() - t0) * 1e3
            best, st = None, None
            for _ in range(3):
                t0 = time.perf_counter(); st = piped(); dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            ingest = {"files": U, "bytes": int(x.nbytes), "where": base or "tmp", "serial_ms": ser_ms, "pipelined_ms": best,
                      "pipelined_gb_per_s": x.nbytes / (best * 1e-3) / 1e9, "threads": int(st.threads), "batches": int(st.batches),
                      "scan_ms": st.scan_s * 1e3, "stage_ms": st.stage_s * 1e3, "read_ms": st.read_s * 1e3,
                      "note": "files in the page cache / tmpfs; wall clock through the Python binding, best of 3"}
        finally:
            shutil.rmtree(tmpd, ignore_errors=True)
    pk = peaks()
    ms_step = tot_ms / args.steps
    value = F * world / (ms_step * 1e-3)
    # per-kernel algorithmic work of one launch over this rank's F frames
    alg = {
        "emis": dict(bytes=F * (4 * D + 4 * N + 4 * G), flops=2.0 * K_AUG * G * F),
        "fwdbwd": dict(bytes=F * (16 * N + 4), flops=0.0),
        "accum": dict(bytes=F * (4 * D + 4 * G + 4 * N), flops=4.0 * K_AUG * G * F),  # two chained contractions
        "mstep": dict(bytes=8.0 * V * api.stats_size(N, M, D), flops=0.0),
    }
    dom = max(kms, key=lambda k: kms[k])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
        traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
    ach = alg[dom]["bytes"] / (kms[dom] * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": traffic, "peak_source": pk["source"], "kernel_ms": kms,
                "tensor_view": {"achieved_tflops": alg[dom]["flops"] / (kms[dom] * 1e-3) / 1e12,
                                "note": "algorithmic flops of the same kernel (2*K*G*F per contraction; 3xTF32 issues three times as many, and G = 80 occupies 80 of 128 MMA rows)"}}
    out = {
        "metric": "frames/sec per Baum-Welch EM iteration", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "frames_per_gpu": F, "utterances_per_gpu": U, "words": V,
                   "l2": "flushed between timed iterations (256 MiB write)", "cpu_affinity": numa, "parallelism": ("utterances sharded, 1 all-reduce of statistics per iteration (%s)" % ar_kind) if world > 1 else "single GPU"},
        "e2e": {"value": F * world / (e2e_ms / e2e_steps * 1e-3), "unit": "frames/s", "ms_per_step": e2e_ms / e2e_steps,
                "h2d_bytes_per_step": int(x.nbytes + off.nbytes),
                "d2h_bytes_per_step": int(8 * (3 * V + 1))},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "decode": dec, "init_model": init, "ingest": ingest,
        "wall_s_timed_region": wall,
    }
    ctx.close()
    del xdev, flush
    if world == 1 and not args.no_regimes:
        out["regimes"] = run_regimes(api, local, pk, tf32_peak_tflops(torch.device("cuda", local)))
    if rank == 0:
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(w, budget_s=args.cpu_seconds)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def tf32_peak_tflops(device):
    """Dense TF32 throughput of this GPU by the protocol of MEASURED_PEAKS.json (library GEMM 8192^3,
    best of 10, CUDA events): the denominator of the tensor-pipe roofline.  Measurement only."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=device)
        b = torch.randn(8192, 8192, device=device)
        best = 1e9
        for _ in range(3):
            torch.matmul(a, b)
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


REGIMES = {
    # bounded slices of BASELINE configs[3] and configs[4]: the decode / large-mixture regimes where the
    # emission contraction is tensor-pipe bound and the cell scorers are HBM bound
    "c4_slice": dict(V=1000, N=5, M=3, U=200, desc="configs[3] slice: 1,000-word model set (N=5, M=3), 200 of 100k utterances"),
    "c5_slice": dict(V=200, N=3, M=128, U=200, desc="configs[4] slice: N=3, M=128, 200 of 2,000 models, 200 utterances"),
}


def run_regimes(api, local, pk, tf32_peak):
    """Decode legs at the shapes where the north-star roofline targets apply (device-resident features):
    emissions of every frame against the whole model set (tensor pipe), forward / Viterbi cell scorers (HBM)."""
    import torch
    from speech_recognition_hmm_continuous_b200 import synth
    out = {}
    for name, w in REGIMES.items():
        V, N, M, U = w["V"], w["N"], w["M"], w["U"]
        cen, s = synth.make_centres(V, N, M, D, seed=77)
        labels = (np.arange(U) % V).astype(np.int32)
        x, off = synth.make_utterances(cen, s, labels, seed=78)
        F = int(off[-1])
        ctx = api.Context(local, timing=True)
        xdev = torch.from_numpy(x).to(torch.device("cuda", local))
        ctx.set_features_device(xdev.data_ptr(), off, D)
        ctx.set_models(api.ModelSet.from_dict(synth.make_models(cen, s)))
        res = {"workload": w["desc"], "frames": F, "models": V, "gaussians_per_frame": V * N * M}
        for leg, fn in (("forward", ctx.forward_scores), ("viterbi", ctx.viterbi_scores)):
            fn()
            em, sc = [], []
            for _ in range(3):
                sco = fn()
                em.append(ctx.kernel_ms("emis"))
                sc.append(ctx.kernel_ms("score" if leg == "forward" else "viterbi"))
            em_ms, sc_ms = float(np.median(em)), float(np.median(sc))
            flops = 2.0 * K_AUG * V * N * M * F               # algorithmic; the 3xTF32 scheme issues three times as many
            sbytes = 4.0 * N * V * F + (N * V * F if leg == "viterbi" else 0.0)   # 4N (+N) bytes per (frame, model)
            res[leg] = {
                "emis_ms": em_ms, "score_ms": sc_ms,
                "emis_tflops_algorithmic": flops / (em_ms * 1e-3) / 1e12,
                "emis_frac_tf32_peak_algorithmic": flops / (em_ms * 1e-3) / 1e12 / tf32_peak,
                "emis_frac_tf32_peak_issued_3x": 3.0 * flops / (em_ms * 1e-3) / 1e12 / tf32_peak,
                "score_gbs": sbytes / (sc_ms * 1e-3) / 1e9, "score_frac_hbm": sbytes / (sc_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "frames_per_s": F / ((em_ms + sc_ms) * 1e-3), "frame_model_pairs_per_s": F * V / ((em_ms + sc_ms) * 1e-3),
            }
        lab, _ = ctx.rank(sco)
        res["top1_matches_generating_word"] = float(np.mean(lab == labels))
        out[name] = res
        ctx.close()
        del xdev
    out["tf32_peak_tflops"] = tf32_peak
    out["tf32_peak_note"] = "library TF32 GEMM 8192^3 measured in this run (same protocol as MEASURED_PEAKS.json); 3xTF32 issues 3 MMAs per algorithmic MMA"
    return out


# ========================================================================== reference arm ====
def _ref_tag(M):
    return "d39m16" if M <= 16 else "d39m128"


def _write_sample(tmp, w, n_utt, word, seed):
    """Feature files + list file for `n_utt` utterances of one word; returns (list path, frames)."""
    from oracle import ref as r
    from speech_recognition_hmm_continuous_b200 import synth
    cen, s = synth.make_centres(w["V"], w["N"], w["M"], D, seed=1234)
    x, off = synth.make_utterances(cen, s, [word] * n_utt, seed=seed)
    files = []
    for u in range(n_utt):
        f = os.path.join(tmp, "w%d_s%d_u%d.bin" % (word, seed, u))
        r.write_features(f, x[off[u]:off[u + 1]])
        files.append(f)
    lst = os.path.join(tmp, "list_w%d_s%d.txt" % (word, seed))
    open(lst, "w").write("\n".join(files) + "\n")
    return lst, int(off[-1])


def _ref_train_once(tag, N, M, lst, out_hmm):
    """Runs the reference trainer binary; returns (seconds excluding the initial-model builder, iterations)."""
    from oracle import ref as r
    t0 = time.perf_counter()
    r.run_train_cli(tag, "w", N, M, lst, out_hmm, stack_unlimited=(M > 16))
    t_total = time.perf_counter() - t0
    _, its = r.parse_train_report(out_hmm[:-4] + ".txt")
    t1 = time.perf_counter()
    r.RefTrain(tag).init_model(N, M, lst)  # creating_initial_model alone (T-FS:732), same files
    t_init = time.perf_counter() - t1
    return max(t_total - t_init, 1e-6), its


def _ref_worker(q, tag, N, M, lst, out_hmm):
    q.put(_ref_train_once(tag, N, M, lst, out_hmm))


def ref_parallel_pass(w, cores, n_utt, tmp, seed):
    """One bounded sample: `cores` reference trainer processes side by side, one word model each."""
    import multiprocessing as mp
    tag = _ref_tag(w["M"])
    jobs = [_write_sample(tmp, w, n_utt, k % w["V"], seed + k) for k in range(cores)]
    q = mp.Queue()
    procs = [mp.Process(target=_ref_worker, args=(q, tag, w["N"], w["M"], jobs[k][0], os.path.join(tmp, "m%d_%d.hmm" % (seed, k)))) for k in range(cores)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.perf_counter() - t0
    frame_iters = sum(jobs[k][1] for k in range(cores)) * np.mean([r[1] for r in res])
    t_em = max(r[0] for r in res)
    return frame_iters, t_em, wall


def cpu_baseline(w, budget_s=15.0):
    """The reference trainer (oracle/_ref, gcc -O2, single-threaded by construction) on one host core,
    on a bounded sample of the same workload; initial-model time excluded."""
    from oracle import ref as r
    tag = _ref_tag(w["M"])
    if not r.available(tag):
        return {"value": None, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref missing"}
    tmp = tempfile.mkdtemp()
    n_utt = 12
    lst, frames = _write_sample(tmp, w, n_utt, 0, seed=4321)
    t, its = _ref_train_once(tag, w["N"], w["M"], lst, os.path.join(tmp, "m.hmm"))
    reps = int(max(0, min(8, budget_s / max(t, 1e-3) - 1)))
    ts = [t] + [_ref_train_once(tag, w["N"], w["M"], lst, os.path.join(tmp, "m.hmm"))[0] for _ in range(reps)]
    t = float(np.median(ts))
    return {"value": frames * its / t, "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": "%d utterances (%d frames) of one word, %d EM iterations, reference trainer binary (oracle/_ref/hmm_fs_%s), init excluded, median of %d runs"
                      % (n_utt, frames, its, tag, len(ts)), "host_cores_available": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref as r
    w = WORKLOADS[args.workload]
    tag = _ref_tag(w["M"])
    if not r.available(tag):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/hmm_fs_%s not built (oracle/build_ref.sh needs /root/reference)" % tag}))
        return
    cores = os.cpu_count() or 1
    tmp = tempfile.mkdtemp()
    n_utt = 6
    for k in range(args.warmup and 1):
        ref_parallel_pass(w, cores, n_utt, tmp, seed=100 + k)
    fi, tem, wall_all = 0.0, 0.0, 0.0
    steps = max(1, min(args.steps, 5))
    for k in range(steps):
        a, b, c = ref_parallel_pass(w, cores, n_utt, tmp, seed=200 + k)
        fi += a
        tem += b
        wall_all += c
    value = fi / tem
    sample = "%d processes x %d utterances each (one word model per process), reference trainer binary, init excluded" % (cores, n_utt)
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec per Baum-Welch EM iteration", "value": value, "unit": "frames/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": 1, "ms_per_step": tem / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"]},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-regimes", action="store_true", help="skip the decode-regime legs (c4 / c5 slices)")
    ap.add_argument("--no-ingest", action="store_true", help="skip the feature-file ingest leg")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to its GPU's NUMA-local CPUs")
    ap.add_argument("--torch-allreduce", action="store_true", help="all-reduce through torch.distributed instead of the NCCL C API")
    ap.add_argument("--upload-chunks", type=int, default=0, help="override the library's upload chunk count (experiments)")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=INT", help="hmmcu_set_option switches (experiments)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
