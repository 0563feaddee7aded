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

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        t0, t1 = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        for stamp, line in self.lines:
            if stamp < t0 or stamp > t1 + 0.03:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9 or parts[0] != str(self.device_index):
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    threads; one step = every reference model object (one per stream, as the reference is used) processes one block."""
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
        secs = O.ref_bench_steps(path, frames, T, ipt, args.warmup, args.steps, quality=quality)
        units = T * ipt * frames * args.steps
        value = units / secs
    line = {
        "impl": "reference", "metric": "audio samples/sec (batch x buffer)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TITLES[args.workload], "model": source, "frames": frames,
                   "streams_per_step": T * ipt, "note": "bounded sample: one reference model object per stream, %d per host thread" % ipt},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": T, "kind": "reference",
                         "sample": "%d threads x %d model objects x %d-frame Process() calls x %d steps, white noise" % (T, ipt, frames, args.steps),
                         "isa": os.path.basename(O.ref_lib_path())},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args, rank, local_rank, world):
    import numpy as np
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

    # ---- load: rank 0 reads the file; ONE NCCL broadcast carries the weights to the other ranks --------------------
    ext = os.path.splitext(path)[1]
    if world > 1:
        if rank == 0:
            blob = torch.frombuffer(bytearray(open(path, "rb").read()), dtype=torch.uint8).to(dev)
            size = torch.tensor([blob.numel()], dtype=torch.int64, device=dev)
        else:
            size = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.broadcast(size, src=0)
        if rank != 0:
            blob = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
        dist.broadcast(blob, src=0)
        data = bytes(blob.cpu().numpy().tobytes())
    else:
        data = open(path, "rb").read()
    loader = na.NeuralModelLoader()
    loader.SetDevice(local_rank)
    loader.SetDefaultQualityScaleFactor(quality)
    loader.SetDefaultNumStreams(streams)     # this rank's shard of the stream batch (contiguous block, no exchange step)
    model = loader.CreateFromMemory(data, ext)

    stream = torch.cuda.ExternalStream(model.GetCudaStream(), device=dev)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nbuf = 4
    xs = [(torch.rand((streams, frames), generator=g, dtype=torch.float32) * 2 - 1).to(dev) for _ in range(nbuf)]
    ys = [torch.empty((streams, frames), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident hot path -------------------------------------------------------------------------------
    # (the clock sampler was started before the model was loaded: nvidia-smi needs ~1 s before its first sample)
    t_wait = time.time()
    while sampler.proc is not None and not sampler.lines and time.time() - t_wait < 2.0:
        time.sleep(0.05)
    for i in range(max(args.warmup, 3)):
        model.ProcessBatch(xs[i % nbuf], ys[i % nbuf], streams, frames)
    model.Synchronize()
    barrier()
    sampler.mark_begin()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        model.ProcessBatch(xs[i % nbuf], ys[i % nbuf], streams, frames)
    e1.record(stream)
    e1.synchronize()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches_per_step = 1 if args.workload.startswith("lstm") else -(-frames // (256 if args.workload in ("a2_full", "a1_nano") else 128))
    if args.workload == "a1_standard" and os.environ.get("NAB200_USE_TC", "2") == "2" and os.environ.get("NAB200_TS_SPLIT", "0") != "0":
        launches_per_step *= 2   # TS kernel, split launch: one kernel per layer array

    # ---- end to end through the public C-ABI call with HOST buffers (pinned), copies inside the timed region ----
    # (a) blocking NA_ProcessBatch, like the reference's Process: H2D + kernel + D2H + wait, one call at a time;
    # (b) NA_ProcessBatchAsync + NA_WaitBatches: the same per-step work, consecutive calls pipelined (the copies of one call
    #     overlap the kernels of its neighbours); every step's result is read back on the host inside the timed region.
    xh = [torch.empty((streams, frames), dtype=torch.float32).pin_memory() for _ in range(3)]
    yh = [torch.empty((streams, frames), dtype=torch.float32).pin_memory() for _ in range(3)]
    for t in xh:
        t.copy_(xs[0].cpu())
    for i in range(3):
        model.ProcessBatch(xh[i % 3], yh[i % 3], streams, frames)
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(args.steps):
        model.ProcessBatch(xh[i % 3], yh[i % 3], streams, frames)
        checksum += float(yh[i % 3][0, 0])
    e2e_blocking_s = time.perf_counter() - t0
    barrier()
    for i in range(3):
        model.ProcessBatchAsync(xh[i % 3], yh[i % 3], streams, frames)
    model.WaitBatches(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        model.ProcessBatchAsync(xh[i % 3], yh[i % 3], streams, frames)
        if i >= 1:
            model.WaitBatches(1)                       # step i-1 is complete: read its result on the host
            checksum += float(yh[(i - 1) % 3][0, 0])
    model.WaitBatches(0)
    checksum += float(yh[(args.steps - 1) % 3][0, 0])
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()   # sampled every 20 ms; only samples from the start of the device-timed loop to the end of the end-to-end loops count

    times = torch.tensor([dev_ms, e2e_s * 1e3, e2e_blocking_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_blocking_ms = float(times[0]), float(times[1]), float(times[2])

    if rank == 0:
        units = world * streams * frames * args.steps
        value = units / (dev_ms * 1e-3)
        alg = algorithmic_bytes_per_stream_call(path, frames, quality)
        peak, peak_src = measured_peak_gbs()
        # the hot path of one step is `launches_per_step` back-to-back launches of our kernels; the roofline is taken over the
        # step (algorithmic bytes of the step / device time of the step)
        kernel_s = dev_ms * 1e-3 / args.steps
        achieved = alg["total"] * streams / kernel_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "audio samples/sec (batch x buffer)", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TITLES[args.workload], "model": source, "streams_per_gpu": streams, "frames": frames,
                       "input": "white noise U[-1,1), seed 1234+rank, %d rotating device buffers" % nbuf,
                       "l2": "per-step working set (stream state) %.0f MB per GPU >> 126 MB L2; no flush needed" % (alg["total"] * streams / 1e6),
                       "sharding": "contiguous stream blocks per rank, one NCCL broadcast of the model at load, no per-step collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_stream_call": alg,
                         "frac_read_only": (alg["read"] * streams / kernel_s / 1e9) / peak,
                         "kernel_us": kernel_s * 1e6},
            "e2e": {"value": world * streams * frames * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": streams * frames * 4, "d2h_bytes_per_step": streams * frames * 4,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "NA_ProcessBatchAsync + NA_WaitBatches with pinned host buffers: every step copies its input in and its output "
                           "out and the host reads each result; consecutive steps are pipelined (two in flight)",
                    "blocking_value": world * streams * frames * args.steps / (e2e_blocking_ms * 1e-3),
                    "blocking_ms_per_step": e2e_blocking_ms / args.steps,
                    "blocking_api": "NA_ProcessBatch with pinned host pointers (H2D + kernel + D2H + wait per call)"},
            "gpu_launches": args.steps * launches_per_step,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(path, frames, quality, args.cpu_seconds)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    tmp.cleanup()


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
    ap.add_argument("--ref-instances-per-thread", type=int, default=4)
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
