#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json: audio samples/s of the batched Process() hot path.

    python bench.py --gpus 1 --steps K --warmup W                      # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus 1 --steps K --warmup W     # the reference's own CPU Process() on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

A "step" is ONE call of the hot path over one batch: `streams` independent mono streams advance by `frames` samples
(default workload = BASELINE.json configs[1]: NAM A1 WaveNet 'Standard', 4096 streams x 128 frames per GPU).
Prints one JSON line (rank 0).  See DESIGN.md section "Measurement" for the byte model behind `roofline`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (fixture file, synthetic golden with the same architecture, streams per GPU, frames per call, quality)
    "a1_standard": ("BossWN-standard.nam", "syn_a1_standard", 4096, 128, 1.0),
    "a1_nano": ("BossWN-nano.nam", "syn_a1_nano", 4096, 128, 1.0),
    "lstm_1x16": ("BossLSTM-1x16.nam", "syn_lstm_1x16", 8192, 128, 1.0),
    "a2_full": ("BossWN-a2.nam", "syn_a2_full", 4096, 256, 1.0),
}
WORKLOAD_TITLES = {
    "a1_standard": "NAM A1 WaveNet 'Standard', batch 4096 streams, buffer 128, per B200",
    "a1_nano": "NAM A1 WaveNet 'Nano', batch 4096 streams, buffer 128, per B200",
    "lstm_1x16": "NAM LSTM 1x16, batch 8192 streams, buffer 128, per B200",
    "a2_full": "NAM A2 'Full' composite (slimmable), quality 1.0, batch 4096, buffer 256, per B200",
}


def model_file(workload, tmpdir):
    """The reference's own fixture when it is staged (oracle/_ref/models, git-ignored), else a synthetic model of
    the same architecture with seeded random weights (committed under tests/golden)."""
    import numpy as np
    fixture, syn, _s, _n, _q = WORKLOADS[workload]
    p = os.path.join(ROOT, "oracle", "_ref", "models", fixture)
    if os.path.exists(p):
        return p, "fixture:" + fixture
    z = np.load(os.path.join(ROOT, "tests", "golden", syn + ".npz"))
    model = json.loads(str(z["model"]))
    model["weights"] = [float(v) for v in z["weights"]]
    p = os.path.join(tmpdir, syn + ".nam")
    with open(p, "w") as f:
        json.dump(model, f)
    return p, "synthetic:" + syn


def algorithmic_bytes_per_stream_call(path, frames, quality=1.0):
    """SURVEY.md section 8d byte model, per stream per call of N frames, state resident in HBM, weights on chip:
    history read  = sum over convs of C * |union_j ([-j*d, N-1-j*d] intersected with negatives)| * 4
    history write = sum over convs of C * min(N, (K-1)*d) * 4
    I/O           = 4 N in + 4 N out.      LSTM: (h, c) read + written once per call."""
    with open(path) as f:
        mj = json.load(f)
    if path.endswith(".nam") and mj.get("architecture") == "SlimmableContainer":
        subs = sorted(mj["config"]["submodels"], key=lambda s: s["max_value"])
        pick = subs[-1]
        for s in subs:
            pick = s
            if quality <= s["max_value"]:
                break
        mj = pick["model"]
    N = frames
    if not path.endswith(".nam") or mj.get("architecture") == "LSTM":
        if path.endswith(".nam"):
            H, L = mj["config"]["hidden_size"], mj["config"]["num_layers"]
        else:
            H, L = mj["layers"][0]["shape"][-1], len(mj["layers"]) - 1
        state = 2 * H * L * 4
        return dict(read=state + 4 * N, write=state + 4 * N, total=2 * state + 8 * N)
    convs = []
    for lc in mj["config"]["layers"]:
        C = lc["channels"]
        ks = lc["kernel_sizes"] if "kernel_sizes" in lc else [lc["kernel_size"]] * len(lc["dilations"])
        for k, d in zip(ks, lc["dilations"]):
            convs.append((C, k, d))
        if "head" in lc and isinstance(lc["head"], dict) and lc["head"].get("kernel_size", 1) > 1:
            convs.append((C, lc["head"]["kernel_size"], 1))
    rd = wr = 0
    for C, K, d in convs:
        cols = set()
        for j in range(1, K):
            lo, hi = -j * d, min(N - 1 - j * d, -1)
            if hi >= lo:
                if hi - lo > 4096:
                    cols.update(range(lo, hi + 1))
                else:
                    cols.update(range(lo, hi + 1))
        rd += C * len(cols) * 4
        wr += C * min(N, (K - 1) * d) * 4
    return dict(read=rd + 4 * N, write=wr + 4 * N, total=rd + wr + 8 * N)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvidia-smi needs ~1 s before its first sample, so it is started before the warm-up; every line is stamped with the host
    # clock on arrival and only the lines that arrived inside [mark_begin, mark_end] (the timed loops) are used

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", os.environ.get("NAB200_BENCH_SMI_MS", "50")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self, name):
        if not hasattr(self, "marks"):
            self.marks = {}
        self.marks[name] = time.time()

    def stop(self):
        if not self.proc:
            return {"burst": {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        marks = getattr(self, "marks", {})

        def window(t0, t1):
            sm, smax, power, reasons = [], [], [], set()
            for stamp, line in self.lines:
                if stamp < t0 or stamp > t1 + 0.03:
                    continue
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9 or parts[0] != str(self.device_index):
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                    power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                    "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}

        # "burst": from the start of the device-timed loop to the end of the end-to-end loops; "sustained": the >= 2 s leg
        out = {"burst": window(marks.get("burst_begin", 0.0), marks.get("sustained_begin", marks.get("burst_end", float("inf"))))}
        if "sustained_begin" in marks:
            out["sustained"] = window(marks["sustained_begin"], marks.get("sustained_end", float("inf")))
        return out


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy kernel read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU Process() (oracle/_ref, compiled from /root/reference) on all host
    threads, 8 model objects per thread (one per stream, as the reference is used - the same instance count as the
    cpu_baseline leg).  One step = every object processes `blocks_per_step` consecutive blocks; blocks_per_step is
    calibrated so that the timed region lasts >= --ref-seconds whatever --steps is (a 13 ms region swung 2x between boxes)."""
    if rank != 0:
        return
    from oracle import oracle as O
    with tempfile.TemporaryDirectory() as tmp:
        path, source = model_file(args.workload, tmp)
        _f, _s, streams, frames, quality = WORKLOADS[args.workload]
        T = host_threads()
        ipt = args.ref_instances_per_thread
        if not O.ref_available():
            emit({"impl": "reference", "unavailable": "oracle/_ref/libna_ref.so not built and /root/reference absent"})
            return
        probe = O.ref_bench_steps(path, frames, T, ipt, 2, 8, quality=quality)          # seconds for 8 blocks per object
        per_block = max(probe / 8.0, 1e-6)
        bps = max(1, int(-(-args.ref_seconds // (per_block * args.steps))))
        secs = O.ref_bench_steps(path, frames, T, ipt, args.warmup * bps, args.steps * bps, quality=quality)
        units = T * ipt * frames * args.steps * bps
        value = units / secs
    line = {
        "impl": "reference", "metric": "audio samples/sec (batch x buffer)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TITLES[args.workload], "model": source, "frames": frames,
                   "streams_per_step": T * ipt, "blocks_per_step": bps, "timed_seconds": secs,
                   "note": "bounded sample: one reference model object per stream, %d per host thread; a step = %d consecutive %d-frame "
                           "Process() calls of every object (sized so the timed region lasts >= %.1f s)" % (ipt, bps, frames, args.ref_seconds)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": T, "kind": "reference",
                         "sample": "%d threads x %d model objects x %d-frame Process() calls x %d steps x %d blocks, white noise, %.2f s timed" % (T, ipt, frames, args.steps, bps, secs),
                         "isa": os.path.basename(O.ref_lib_path())},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


PARITY_TOL = {"a1_standard": 1e-5, "a1_nano": 1e-5, "a2_full": 1e-5, "lstm_1x16": 5e-5}


def probe_streams(streams, grid_hint=740):
    """Stream slots whose benched output is compared with the oracle: first / middle / last, plus slots a persistent kernel
    serves in its later rounds (last wave)."""
    cand = [0, 1, streams // 3, streams // 2, streams - grid_hint // 2, streams - 2, streams - 1, grid_hint + 5]
    return sorted({c for c in cand if 0 <= c < streams})[:8]


def measure(na, torch, dist, dev, rank, world, workload, path, streams, frames, quality, steps, warmup, sustained_s=0.0, sampler=None,
            amplitude=1.0):
    """Device-timed throughput, parity of the benched outputs against the oracle, end-to-end legs through the C ABI with host
    buffers, and (optionally) a sustained leg, for one workload on this rank's GPU."""
    import numpy as np
    loader = na.NeuralModelLoader()
    loader.SetDevice(dev.index)
    loader.SetDefaultQualityScaleFactor(quality)
    loader.SetDefaultNumStreams(streams)     # this rank's shard of the stream batch (contiguous block, no exchange step)
    model = loader.CreateFromFile(path)
    stream = torch.cuda.ExternalStream(model.GetCudaStream(), device=dev)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nbuf = 4
    xs_host = [((torch.rand((streams, frames), generator=g, dtype=torch.float32) * 2 - 1) * amplitude) for _ in range(nbuf)]
    xs = [t.to(dev) for t in xs_host]
    ys = [torch.empty((streams, frames), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(warmup, 3)
    for i in range(warm):
        model.ProcessBatch(xs[i % nbuf], ys[i % nbuf], streams, frames)
    model.Synchronize()
    barrier()
    if sampler is not None:
        sampler.mark("burst_begin")
    l0 = model.GetKernelLaunchCount()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        model.ProcessBatch(xs[i % nbuf], ys[i % nbuf], streams, frames)
    e1.record(stream)
    e1.synchronize()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = model.GetKernelLaunchCount() - l0
    if sampler is not None:
        sampler.mark("burst_end")

    # ---- parity of what was just timed: replay the call sequence of a few stream slots through the oracle ----------------
    from oracle import oracle as O
    probes = probe_streams(streams)
    last = min(nbuf, steps)
    y_last = {j: ys[j % nbuf][probes].cpu().numpy() for j in range(steps - last, steps)}
    seq = [i % nbuf for i in range(warm)] + [i % nbuf for i in range(steps)]
    worst = 0.0
    for k, sidx in enumerate(probes):
        hist = np.concatenate([xs_host[b][sidx].numpy() for b in seq])
        ref = O.PortModel.from_file(path, quality=quality).process(hist)
        for j in range(steps - last, steps):
            a = (warm + j) * frames
            worst = max(worst, float(np.abs(ref[a:a + frames] - y_last[j][k]).max()))
    tol = PARITY_TOL[workload]
    parity = {"max_abs": worst, "streams_checked": len(probes), "streams": probes, "steps_checked": last, "tol": tol, "ok": bool(worst <= tol),
              "against": "oracle/na_oracle.c (plain-C restatement of the reference, pinned to the compiled reference), replaying the benched calls"}

    # ---- end to end through the public C-ABI call with HOST buffers (pinned), copies inside the timed region ----
    # (a) blocking NA_ProcessBatch, like the reference's Process: the call returns with the output complete;
    # (b) NA_ProcessBatchAsync + NA_WaitBatches: the same per-step work, consecutive calls pipelined (the copies of one call
    #     overlap the kernels of its neighbours); every step's result is read back on the host inside the timed region.
    xh = [torch.empty((streams, frames), dtype=torch.float32).pin_memory() for _ in range(3)]
    yh = [torch.empty((streams, frames), dtype=torch.float32).pin_memory() for _ in range(3)]
    for t in xh:
        t.copy_(xs_host[0])
    for i in range(3):
        model.ProcessBatch(xh[i % 3], yh[i % 3], streams, frames)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(steps):
        model.ProcessBatch(xh[i % 3], yh[i % 3], streams, frames)
        checksum += float(yh[i % 3][0, 0])
    e2e_blocking_s = time.perf_counter() - t0
    barrier()
    for i in range(3):
        model.ProcessBatchAsync(xh[i % 3], yh[i % 3], streams, frames)
    model.WaitBatches(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        model.ProcessBatchAsync(xh[i % 3], yh[i % 3], streams, frames)
        if i >= 1:
            model.WaitBatches(1)                       # step i-1 is complete: read its result on the host
            checksum += float(yh[(i - 1) % 3][0, 0])
    model.WaitBatches(0)
    checksum += float(yh[(steps - 1) % 3][0, 0])
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- sustained: back-to-back steps for >= sustained_s seconds (streaming audio is a sustained workload) -------------
    sus = None
    if sustained_s > 0:
        n_sus = max(steps, int(sustained_s / max(dev_ms * 1e-3 / steps, 1e-6)) + 1)
        if sampler is not None:
            sampler.mark("sustained_begin")
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for i in range(n_sus):
            model.ProcessBatch(xs[i % nbuf], ys[i % nbuf], streams, frames)
        s1.record(stream)
        s1.synchronize()
        barrier()
        if sampler is not None:
            sampler.mark("sustained_end")
        sus = {"steps": n_sus, "ms": s0.elapsed_time(s1)}

    times = [dev_ms, e2e_s * 1e3, e2e_blocking_s * 1e3, sus["ms"] if sus else 0.0]
    tt = torch.tensor(times, dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_blocking_ms, sus_ms = [float(v) for v in tt]
    alg = algorithmic_bytes_per_stream_call(path, frames, quality)
    peak, peak_src = measured_peak_gbs()
    kernel_s = dev_ms * 1e-3 / steps
    achieved = alg["total"] * streams / kernel_s / 1e9
    kernel_name = na.describe_model_file(path).get("kernel") if path.endswith(".nam") else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            ent = json.load(open(tp)).get(workload, {})
            # a capture of another kernel is stale: report nothing rather than a wrong number
            if ent.get("kernel_choice") == kernel_name or kernel_name is None:
                # a step of more than 128 frames is several launches of the tensor-core kernels (128 frames per pass)
                traffic = ent.get("dram_bytes_per_step", ent.get("dram_bytes_per_launch"))
        except Exception:
            traffic = None
    total_units = world * streams * frames
    res = {
        "value": total_units * steps / (dev_ms * 1e-3), "unit": "samples/s", "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warm,
        "streams_per_gpu": streams, "frames": frames, "kernel": kernel_name,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_stream_call": alg,
                     "frac_read_only": (alg["read"] * streams / kernel_s / 1e9) / peak, "kernel_us": kernel_s * 1e6},
        "e2e": {"value": total_units * steps / (e2e_ms * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": streams * frames * 4, "d2h_bytes_per_step": streams * frames * 4,
                "ms_per_step": e2e_ms / steps,
                "api": "NA_ProcessBatchAsync + NA_WaitBatches with pinned host buffers: every step copies its input in and its output "
                       "out and the host reads each result; consecutive steps are pipelined (two in flight)",
                "blocking_value": total_units * steps / (e2e_blocking_ms * 1e-3),
                "blocking_ms_per_step": e2e_blocking_ms / steps,
                "blocking_api": "NA_ProcessBatch with pinned host pointers (the call returns with the output complete; the kernels read and write the page-locked buffers directly)"},
        "parity": parity,
        "gpu_launches": int(launches),
    }
    if sus:
        res["sustained"] = {"value": total_units * sus["steps"] / (sus_ms * 1e-3), "unit": "samples/s", "steps": sus["steps"],
                            "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / sus["steps"],
                            "frac": (alg["total"] * streams / (sus_ms * 1e-3 / sus["steps"]) / 1e9) / peak}
    del model
    return res


def single_stream_latency(na, torch, dev, workload, path, frames, quality, calls=2000):
    """cfg 1 (the reference's own use): ONE stream, one blocking Process() per `frames`-sample host buffer; microseconds per
    call and the real-time factor at 48 kHz, with the reference's CPU figure (1 thread, 1 model object, ModelTest protocol,
    Utils/ModelTest/ModelTest.cpp:59-79) measured beside it on this box, and parity of the processed block sequence."""
    import numpy as np
    from oracle import oracle as O
    loader = na.NeuralModelLoader()
    loader.SetDevice(dev.index)
    loader.SetDefaultQualityScaleFactor(quality)
    loader.SetDefaultNumStreams(1)
    model = loader.CreateFromFile(path)
    rng = np.random.default_rng(4321)
    x = rng.uniform(-1, 1, (64, frames)).astype(np.float32)
    y = np.empty_like(x)
    for i in range(64):
        y[i] = model.Process(x[i])
    ref = O.PortModel.from_file(path, quality=quality).process(x.reshape(-1)).reshape(x.shape)
    worst = float(np.abs(ref - y).max())
    buf = np.ascontiguousarray(x[0])
    t0 = time.perf_counter()
    for i in range(calls):
        model.Process(buf)
    us = (time.perf_counter() - t0) / calls * 1e6
    out = {"us_per_call": us, "value": frames / (us * 1e-6), "unit": "samples/s", "realtime_factor_48k": (frames / 48000.0) / (us * 1e-6),
           "streams": 1, "frames": frames, "calls": calls, "api": "Process(host buffer) through the reference's own 15-function C ABI",
           "parity": {"max_abs": worst, "tol": PARITY_TOL[workload], "ok": bool(worst <= PARITY_TOL[workload]), "blocks": 64}}
    if O.ref_available():
        total, _per = O.ref_bench(path, frames, 2.0, 1, 1, quality)
        out["reference_cpu"] = {"value": total, "unit": "samples/s", "us_per_call": frames / total * 1e6, "cores": 1,
                                "sample": "reference NeuralModel::Process, 1 thread, 1 model object, %d-frame calls for 2 s" % frames}
    del model
    return out


def pin_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: bind this rank's host threads (and with them its page-locked staging buffers, first touched here) to the
    NUMA node its GPU hangs off, so eight ranks do not all stage through one socket.  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            ent = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ent) and ent[local_rank].isdigit():
                idx = int(ent[local_rank])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        sysdev = "/sys/bus/pci/devices/%s:%s" % (dom[-4:], rest)
        node = int(open(os.path.join(sysdev, "numa_node")).read().strip())
        if node < 0:
            return {"numa_node": node, "pinned": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "pinned": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "pinned": True, "cpus": len(cpus)}
    except Exception as e:   # no NVML / no sysfs: run unpinned
        return {"pinned": False, "why": str(e)[:80]}


def run_b200(args, rank, local_rank, world):
    affinity = pin_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import neuralaudio_b200 as na

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    sampler = ClockSampler(local_rank)
    sampler.start()
    _fixture, _syn, streams, frames, quality = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    tmp = tempfile.TemporaryDirectory()
    path, source = model_file(args.workload, tmp.name)
    shard = None
    if world > 1:
        shard = multi_gpu_load_check(na, torch, dist, dev, rank, world, path, quality)
    # (the clock sampler was started before the model was loaded: nvidia-smi needs ~1 s before its first sample)
    t_wait = time.time()
    while sampler.proc is not None and not sampler.lines and time.time() - t_wait < 2.0:
        time.sleep(0.05)
    head = measure(na, torch, dist, dev, rank, world, args.workload, path, streams, frames, quality, args.steps, args.warmup,
                   sustained_s=args.sustained_seconds, sampler=sampler, amplitude=0.5 if args.workload.startswith("lstm") else 1.0)
    clocks = sampler.stop()

    if rank == 0:
        line = {
            "metric": "audio samples/sec (batch x buffer)", "value": head["value"], "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TITLES[args.workload], "model": source, "streams_per_gpu": streams, "frames": frames,
                       "kernel": head["kernel"],
                       "input": "white noise U[-1,1), seed 1234+rank, 4 rotating device buffers",
                       "l2": "per-step working set (stream state) %.0f MB per GPU >> 126 MB L2; no flush needed" % (head["roofline"]["algorithmic_bytes_per_stream_call"]["total"] * streams / 1e6),
                       "sharding": "contiguous stream blocks per rank; one ncclBroadcast of [packed weights | prewarmed state template] issued by the "
                                   "library (NA_BroadcastModel) at load, no per-step collective"},
            "roofline": head["roofline"], "e2e": head["e2e"], "parity": head["parity"], "gpu_launches": head["gpu_launches"],
            "clocks": clocks.get("burst", clocks),
        }
        if "sustained" in head:
            line["sustained"] = dict(head["sustained"], clocks=clocks.get("sustained"))
        if shard is not None:
            line["multi_gpu_load"] = shard
            line["host_affinity"] = {"rank0": affinity, "note": "each rank binds its host threads to its GPU's NUMA node before CUDA starts"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(path, frames, quality, args.cpu_seconds)
        if world == 1 and not args.no_extras and args.workload == "a1_standard":
            # the other BASELINE.json configurations that fit one GPU, each with its own parity probe (cfg 2 is the headline above)
            configs = {}
            for key, wl, st in (("cfg3_lstm_1x16_8192x128", "lstm_1x16", 100), ("cfg5_a2_full_4096x256", "a2_full", 60)):
                p2, src2 = model_file(wl, tmp.name)
                _f, _s, s2, n2, q2 = WORKLOADS[wl]
                r = measure(na, torch, None, dev, 0, 1, wl, p2, s2, n2, q2, st, 5, amplitude=0.5 if wl.startswith("lstm") else 1.0)
                r["workload"] = WORKLOAD_TITLES[wl]
                r["model"] = src2
                configs[key] = r
            p1, src1 = model_file("a1_nano", tmp.name)
            r1 = single_stream_latency(na, torch, dev, "a1_nano", p1, 128, 1.0)
            r1["workload"] = "NAM A1 WaveNet 'Nano', 1 stream, buffer 128 (the reference's own use)"
            r1["model"] = src1
            configs["cfg1_a1_nano_1x128"] = r1
            line["configs"] = configs
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    tmp.cleanup()


def multi_gpu_load_check(na, torch, dist, dev, rank, world, path, quality):
    """Multi-GPU load through the library: rank 0 builds the model (parse, pack, prewarm); the other ranks build the same
    architecture WITHOUT prewarming, and ONE ncclBroadcast issued by the library (NA_BroadcastModel, its own communicator
    bootstrapped from a 128-byte NCCL unique id) carries rank 0's [packed weights | prewarmed state template] to them.
    Returns a small report: bytes broadcast, NCCL ranks, and whether a stream advanced on this rank after the broadcast
    reproduces rank 0's output bit for bit (gathered and compared on rank 0)."""
    import numpy as np
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(na.nccl_get_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)          # bootstrap only: 128 bytes
    comm = na.NcclComm(world, rank, bytes(uid.cpu().numpy().tobytes()), dev.index)
    loader = na.NeuralModelLoader()
    loader.SetDevice(dev.index)
    loader.SetDefaultQualityScaleFactor(quality)
    loader.SetDefaultNumStreams(8)
    model = loader.CreateFromFile(path, doPrewarm=(rank == 0))
    nbytes = model.BroadcastModel(comm, 0)
    rng = np.random.default_rng(99)
    x = rng.uniform(-1, 1, (8, 256)).astype(np.float32)     # the same input on every rank
    y = np.empty_like(x)
    model.ProcessBatch(x, y, 8, 256)
    yt = torch.from_numpy(y).to(dev)
    gathered = [torch.empty_like(yt) for _ in range(world)] if rank == 0 else None
    dist.gather(yt, gathered, dst=0)
    report = None
    if rank == 0:
        same = all(bool(torch.equal(gathered[0], t)) for t in gathered[1:])
        report = {"api": "NA_BroadcastModel (ncclBroadcast inside the library)", "bytes": int(nbytes), "nccl_ranks": comm.nranks,
                  "ranks_bit_identical_to_rank0": same}
    del model
    comm.close()
    return report


def cpu_baseline(path, frames, quality, seconds):
    """The reference's own Process() (oracle/_ref) on the box's host cores: a BOUNDED sample of the same workload."""
    from oracle import oracle as O
    T = host_threads()
    if not O.ref_available():
        # fall back to the plain-C port on one core
        import numpy as np
        m = O.PortModel.from_file(path, quality=quality)
        x = np.random.default_rng(1).uniform(-1, 1, frames).astype(np.float32)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            m.process(x)
            n += frames
        return {"value": n / (time.perf_counter() - t0), "unit": "samples/s", "cores": 1, "kind": "port",
                "sample": "plain-C oracle, 1 stream, %d-frame calls for %.0f s" % (frames, seconds)}
    ipt = 8
    one, _ = O.ref_bench(path, frames, max(2.0, seconds / 4), 1, ipt, quality)
    total, per = O.ref_bench(path, frames, seconds, T, ipt, quality)
    return {"value": total, "unit": "samples/s", "cores": T, "kind": "reference",
            "sample": "reference NeuralModel::Process, %d threads x %d model objects (one per stream), %d-frame calls on white noise for %.0f s" % (T, ipt, frames, seconds),
            "one_thread_value": one, "thread_scaling": total / one if one > 0 else None, "isa": os.path.basename(O.ref_lib_path())}


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: anything a library writes to file descriptor 1 (NCCL prints its version banner
    # there when NCCL_DEBUG is set on the box) is sent to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="a1_standard", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-instances-per-thread", type=int, default=8)
    ap.add_argument("--ref-seconds", type=float, default=2.5, help="--impl reference: minimum length of the timed region")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained leg (0: skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg 1 / 3 / 5 lines of the `configs` map")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
