"""GPU triage for the TS kernel: parity against golden vectors / oracle for issuer counts 1 and 4, then timing."""
import os, sys, time, tempfile, pathlib
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for

def single(name, frames=128):
    g = load_golden(golden_files(name)[0])
    tmp = pathlib.Path(tempfile.mkdtemp())
    mf = model_file_for(g, tmp)
    ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(1)
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
    return y

for tc, iss in ((2, 1),):
    na.set_option("use_tc", tc); na.set_option("ts_issuers", iss)
    print("use_tc", tc, "ts_issuers", iss, flush=True)
    single("ref_BossWN_standard"); single("syn_a1_standard", 37)
    y = batch("ref_BossWN_standard", 600, 128, 40)
    if tc == 2 and iss == 1: y1 = y
    if tc == 2 and iss == 4: print("  issuers 1 vs 4: max diff %.3g" % np.abs(y - y1).max())
