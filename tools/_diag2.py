import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import neuralaudio_b200 as na
from conftest import golden_files, load_golden, model_file_for
import tempfile, pathlib
g = load_golden(golden_files("ref_namcore_wavenet_a1_standard")[0])
mf = model_file_for(g, pathlib.Path(tempfile.mkdtemp()))
for tc in (1,0):
    na.set_option("use_tc", tc)
    m = na.NeuralModelLoader().CreateFromFile(mf)
    x = np.zeros(6144, dtype=np.float32); y = np.empty_like(x)
    for i in range(0, x.size, 128): y[i:i+128] = m.Process(np.ascontiguousarray(x[i:i+128]))
    e = np.abs(y-g["dc"][0]); print("tc",tc,"dc ref",g["dc"][0],"maxabs %.3g at %d"%(e.max(), e.argmax()), "y[:6]", y[:6], "y[500:503]", y[500:503], "y[-3:]", y[-3:])
