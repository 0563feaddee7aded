import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# float tolerances (max-abs against the reference's Internal CPU path, BASELINE.json north_star / SURVEY.md section 7):
WAVENET_TOL = 1e-5     # north_star: "max-abs error <= 1e-5 vs reference CPU"
# the reference's own static and dynamic LSTM builds differ by 5.6e-6..1.1e-5 on identical input;
# an fp32 LSTM is a feedback system, so its tolerance is stated separately:
LSTM_TOL = 5e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files(prefix=None):
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    if prefix:
        files = [f for f in files if os.path.basename(f).startswith(prefix)]
    return files


def golden_id(path):
    return os.path.splitext(os.path.basename(path))[0]


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["name"] = golden_id(path)
    g["info"] = json.loads(str(g["info"]))
    if "model" in g:
        g["model"] = json.loads(str(g["model"]))
    return g


def external_sample_rate_of(g):
    """48 kHz unless the vector was made with the host running faster (OversampleNAMConfig, NeuralModel.cpp:92-130)."""
    return int(g["external_sample_rate"]) if "external_sample_rate" in g else 48000


def is_lstm_case(g):
    name = g["name"].lower()
    return "lstm" in name or "tw40" in name


def tol_for(g):
    return LSTM_TOL if is_lstm_case(g) else WAVENET_TOL


def model_file_for(g, tmp_path):
    """Path of the model file behind a golden vector: synthetic ones are rebuilt from the committed weights,
    reference fixtures come from the staged (git-ignored) oracle/_ref/models; None when not staged."""
    if "model" in g:
        d = dict(g["model"])
        d["weights"] = [float(x) for x in g["weights"]]
        p = os.path.join(str(tmp_path), g["name"] + ".nam")
        with open(p, "w") as f:
            json.dump(d, f)
        return p
    p = os.path.join(ROOT, "oracle", "_ref", "models", str(g["fixture"]))
    return p if os.path.exists(p) else None


@pytest.fixture(scope="session")
def na():
    import neuralaudio_b200
    return neuralaudio_b200
