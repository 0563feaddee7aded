"""GPU triage for the fp16-pair WaveNet kernel (use_tc = 3): parity against golden vectors / the oracle, then timing
beside the 3xTF32 kernel (use_tc = 2)."""
import os, sys, tempfile, pathlib
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for, external_sample_rate_of


def single(name, frames=128):
    g = load_golden(golden_files(name)[0])
    tmp = pathlib.Path(tempfile.mkdtemp())
    mf = model_file_for(g, tmp)
    ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(1); ld.SetExternalSampleRate(external_sample_rate_of(g))
    m = ld.CreateFromFile(mf)
    x = g["x"]; y = np.empty_like(x)
    for i in range(0, x.size, frames):
        y[i:i + frames] = m.Process(np.ascontiguousarray(x[i:i + frames]))
    err = np.abs(y - g["y"])
    print("  single", name, "frames", frames, "maxabs %.3g at %d" % (err.max(), err.argmax()), "first", y[:2], g["y"][:2], flush=True)


def batch(name, streams, frames, calls):
    from oracle import oracle as O
    g = load_golden(golden_files(name)[0])
    tmp = pathlib.Path(tempfile.mkdtemp())
    mf = model_file_for(g, tmp)
    ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(streams)
    m = ld.CreateFromFile(mf)
    rng = np.random.default_rng(5); x = rng.uniform(-1, 1, (calls, streams, frames)).astype(np.float32); y = np.empty_like(x)
    for c in range(calls): m.ProcessBatch(x[c], y[c], streams, frames)
    worst = 0
    for s in sorted(set([0, 1, 7, streams // 2, streams - 1])):
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        worst = max(worst, float(np.abs(ys - y[:, s, :].reshape(-1)).max()))
    print("  batch", name, streams, "x", frames, "x", calls, "worst maxabs %.3g" % worst, flush=True)


def timing(name, tc, ctas, S=4096, n=128, steps=50):
    if "a2_" in name: n = 256
    import torch
    na.set_option("use_tc", tc); na.set_option("h_ctas", ctas)
    g = load_golden(golden_files(name)[0])
    tmp = pathlib.Path(tempfile.mkdtemp())
    mf = model_file_for(g, tmp)
    ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(S)
    m = ld.CreateFromFile(mf)
    x = torch.rand((4, S, n), device="cuda") * 2 - 1
    y = torch.empty_like(x)
    stream = torch.cuda.ExternalStream(m.GetCudaStream())
    for k in range(6): m.ProcessBatch(x[k % 4], y[k % 4], S, n)
    m.Synchronize()
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps): m.ProcessBatch(x[k % 4], y[k % 4], S, n)
        e1.record(stream); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    print("  timing %s use_tc=%d ctas/SM=%d: %.1f us/call  %.3f Gsamples/s" % (name, tc, ctas, best * 1e3, S * n / best / 1e6), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        na.set_option("use_tc", 3)
        print("use_tc 3", na.describe_model_file(model_file_for(load_golden(golden_files("syn_a1_standard.")[0]), pathlib.Path(tempfile.mkdtemp())), 48000).get("kernel"), flush=True)
        single("ref_BossWN_standard"); single("syn_a1_standard.", 37); single("syn_a1_lite", 64); single("syn_a1_standard_sr96000", 128)
        single("ref_namcore_wavenet_a1_standard", 1 + 127)
        batch("ref_BossWN_standard", 900, 128, 40)
    if what in ("all", "parity", "a2"):
        na.set_option("use_tc", 3)
        single("syn_a2_full", 256); single("ref_BossWN_a2.", 128); single("ref_BossWN_a2_q0", 100); single("syn_a2_lite", 7)
        batch("syn_a2_full", 700, 256, 12)
    if what in ("all", "a2"):
        for tc, ctas in ((0, 0), (3, 0), (3, 3)):
            timing("syn_a2_full", tc, ctas)
        for tc, ctas in ((0, 0), (3, 0)):
            timing("syn_a2_lite", tc, ctas)
    if what == "timing_a2":
        timing("syn_a2_full", 3, 0, steps=8)
    if what == "timing1":
        timing("syn_a1_standard.", 3, int(os.environ.get("NAB200_H_CTAS", "0")), steps=8)
    if what in ("all", "timing"):
        for tc, ctas in ((2, 0), (3, 4), (3, 5), (3, 3)):
            timing("syn_a1_standard.", tc, ctas)
