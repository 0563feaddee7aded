"""Max-abs error of the CUDA path against every committed golden vector (outputs of the unmodified reference), one line per
vector: the numbers behind the tolerances in tests/conftest.py.  Usage: python tools/parity_report.py > profiles/<round>_parity.txt"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as C
import neuralaudio_b200 as na


def main():
    worst = {"wavenet": 0.0, "lstm": 0.0}
    with tempfile.TemporaryDirectory() as tmp:
        for path in C.golden_files():
            g = C.load_golden(path)
            mf = C.model_file_for(g, tmp)
            if mf is None:
                print("%-40s fixture not staged" % g["name"])
                continue
            ld = na.NeuralModelLoader()
            ld.SetExternalSampleRate(C.external_sample_rate_of(g))
            ld.SetDefaultQualityScaleFactor(float(g.get("quality", 1.0)))
            m = ld.CreateFromFile(mf)
            x = g["x"]
            y = np.empty_like(x)
            for i in range(0, x.size, 128):
                y[i:i + 128] = m.Process(np.ascontiguousarray(x[i:i + 128]))
            err = float(np.abs(y - g["y"]).max())
            kind = "lstm" if C.is_lstm_case(g) else "wavenet"
            worst[kind] = max(worst[kind], err)
            d = na.describe_model_file(mf, C.external_sample_rate_of(g))
            kernel = d.get("kernel") or (d.get("submodels") or [{}])[-1].get("model", {}).get("kernel", "")
            print("%-40s max-abs %.3g   |y|max %.3f   tol %.0e   %s" % (g["name"], err, float(np.abs(g["y"]).max()), C.tol_for(g), kernel))
    print("worst: WaveNet %.3g (tolerance %.0e), LSTM %.3g (tolerance %.0e)" % (worst["wavenet"], C.WAVENET_TOL, worst["lstm"], C.LSTM_TOL))


if __name__ == "__main__":
    main()
