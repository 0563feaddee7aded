"""Small workload over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_run.py [names ...]"""
import os
import pathlib
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for, external_sample_rate_of

DEFAULT = ("syn_a1_standard.", "syn_a1_nano.", "syn_a2_full", "syn_dyn_20x10", "syn_lstm_1x16", "syn_lstm_2x8", "syn_dyn_lstm_3x18")
for name in (sys.argv[1:] or DEFAULT):
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, pathlib.Path(tempfile.mkdtemp()))
    S = 70
    ld = na.NeuralModelLoader()
    ld.SetExternalSampleRate(external_sample_rate_of(g))
    ld.SetDefaultNumStreams(S)
    m = ld.CreateFromFile(mf)
    x = np.random.default_rng(1).uniform(-0.5, 0.5, (3, S, 100)).astype(np.float32)
    y = np.empty_like(x)
    for k in range(3):
        m.ProcessBatch(x[k], y[k], S, 100)
    # the single-stream path too (small WaveNets take the one-CTA kernel)
    y1 = m.Process(np.ascontiguousarray(x[0, 0]))
    print(name, "ok", float(np.abs(y).max()), float(np.abs(y1).max()))
