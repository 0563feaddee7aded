import os, sys, numpy as np, tempfile, pathlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for
for name in ("syn_a1_standard", "syn_dyn_20x10"):
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, pathlib.Path(tempfile.mkdtemp()))
    ld = na.NeuralModelLoader(); ld.SetDefaultNumStreams(700)
    m = ld.CreateFromFile(mf)
    x = np.random.default_rng(1).uniform(-1, 1, (3, 700, 100)).astype(np.float32); y = np.empty_like(x)
    for k in range(3): m.ProcessBatch(x[k], y[k], 700, 100)
    print(name, "ok", float(np.abs(y).max()))
