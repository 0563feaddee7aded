"""First-run triage on the GPU box: each stage runs in its own subprocess under a timeout so that a hang or crash in
one kernel does not hide the others.  Prints numbers, asserts nothing.  Usage: python tools/gpu_diag.py"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

STAGE = r'''
import sys, os, json, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for
import tempfile, pathlib
name, tma, streams, frames, calls = %(name)r, %(tma)d, %(streams)d, %(frames)d, %(calls)d
na.set_option("use_tma", tma)
g = load_golden(golden_files(name)[0])
tmp = pathlib.Path(tempfile.mkdtemp())
mf = model_file_for(g, tmp)
ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(streams)
m = ld.CreateFromFile(mf)
if streams == 1:
    x = g["x"][:frames*calls]; y = np.empty_like(x)
    for i in range(0, x.size, frames): y[i:i+frames] = m.Process(np.ascontiguousarray(x[i:i+frames]))
    err = np.abs(y - g["y"][:x.size]); print("RESULT", name, "tma", tma, "single", frames, "maxabs %%.3g at %%d" %% (err.max(), err.argmax()), "first", y[:3], g["y"][:3])
else:
    from oracle import oracle as O
    rng = np.random.default_rng(5); x = rng.uniform(-1,1,(calls,streams,frames)).astype(np.float32); y = np.empty_like(x)
    for c in range(calls): m.ProcessBatch(x[c], y[c], streams, frames)
    worst = 0
    for s in sorted(set([0, 1, 7, 8, streams//2, streams-1])):
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        e = float(np.abs(ys - y[:, s, :].reshape(-1)).max()); worst = max(worst, e)
    print("RESULT", name, "tma", tma, "batch", streams, "x", frames, "worst maxabs %%.3g" %% worst)
'''


def run(name, tma, streams, frames, calls):
    code = STAGE % dict(root=ROOT, name=name, tma=tma, streams=streams, frames=frames, calls=calls)
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
        out = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        print("\n".join(out) if out else "FAILED %s tma=%d S=%d n=%d rc=%d\n%s" % (name, tma, streams, frames, p.returncode, (p.stderr or p.stdout)[-1500:]))
    except subprocess.TimeoutExpired:
        print("TIMEOUT %s tma=%d S=%d n=%d" % (name, tma, streams, frames))
    sys.stdout.flush()


def tc_first():
    for name in ["syn_a1_standard", "syn_a1_lite", "ref_BossWN_standard"]:
        run(name, 1, 1, 128, 16)
        run(name, 1, 1, 37, 16)
    run("syn_a1_standard", 1, 40, 128, 8)
    run("syn_a1_standard", 1, 700, 128, 3)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tc":
        tc_first()
        sys.exit(0)
    subprocess.run(["nvidia-smi", "--query-gpu=name,driver_version,memory.total", "--format=csv"])
    for name in ["syn_lstm_1x16", "syn_lstm_2x8", "ref_tw40"]:
        run(name, 0, 1, 128, 16)
    for tma in (0, 1):
        for name in ["syn_a1_nano", "syn_a1_feather", "syn_a1_lite", "syn_a1_standard", "syn_a2_lite", "syn_a2_full"]:
            run(name, tma, 1, 128, 16)
            run(name, tma, 1, 32, 16)
        run("syn_a1_standard", tma, 40, 128, 8)
        run("syn_a2_full", tma, 24, 256, 4)
    run("syn_lstm_1x16", 0, 70, 128, 4)
