"""CPU: pins the oracle.  The plain-C restatement (oracle/na_oracle.c) must reproduce every committed golden vector
(outputs of the unmodified reference, tests/golden/make_golden.py) and, where the compiled reference is staged
(oracle/_ref), the reference itself on fresh inputs."""
import json
import os

import numpy as np
import pytest

from conftest import golden_files, golden_id, load_golden, model_file_for, tol_for, is_lstm_case, external_sample_rate_of
from oracle import oracle as O


@pytest.mark.parametrize("path", golden_files(), ids=golden_id)
def test_port_reproduces_golden(path, tmp_path):
    g = load_golden(path)
    mf = model_file_for(g, tmp_path)
    if mf is None:
        pytest.skip("fixture model not staged (oracle/_ref/models)")
    q = float(g.get("quality", 1.0))
    sr = external_sample_rate_of(g)
    m = O.PortModel.from_file(mf, quality=q, external_sample_rate=sr)
    y = m.process(g["x"])
    err = float(np.abs(y - g["y"]).max())
    assert err <= tol_for(g), "port vs golden max-abs %.3g" % err
    m2 = O.PortModel.from_file(mf, quality=q, external_sample_rate=sr)
    dc = m2.process(np.zeros(512, dtype=np.float32))
    assert float(np.abs(dc - g["dc"]).max()) <= tol_for(g)
    if not is_lstm_case(g):
        # prewarm leaves a WaveNet at the silence fixed point: constant output
        assert float(np.abs(dc - dc[0]).max()) <= 1e-6
        assert m2.receptive_field() == (g["info"]["rf"] if g["info"]["static"] else m2.receptive_field())


# known answers of the reference (x[i] = sin(0.01 i), 128-sample calls, default loader)
KNOWN = {
    "BossWN-nano.nam": (1.0, 0.000353399199, -0.252702147, -0.283021569, 21.8273232),
    "BossWN-feather.nam": (1.0, -0.00016338109, -0.2728616, -0.295489967, 28.6872998),
    "BossWN-standard.nam": (1.0, -0.00067000452, -0.317638844, -0.349413633, 39.1764422),
    "BossLSTM-1x16.nam": (1.0, -0.0270614624, -0.253457189, -0.171944439, -219.212568),
    "BossLSTM-2x8.nam": (1.0, 0.00714398921, -0.170382544, -0.184256151, -172.00273),
    "tw40_blues_deluxe_deerinkstudios.json": (1.0, 0.00412131473, -0.179074824, -0.300372869, 28.8511792),
}
KNOWN_A2 = {1.0: (0.000276284292, -0.257261902, -0.284163564, 30.4085193), 0.0: (0.000122590631, -0.236403778, -0.282012135, 14.3841253)}


def _sine():
    return np.sin(np.arange(4096, dtype=np.float64) * 0.01).astype(np.float32)


@pytest.mark.skipif(not O.ref_available(), reason="compiled reference not staged")
@pytest.mark.parametrize("name", sorted(KNOWN))
def test_compiled_reference_known_answers(name):
    p = O.model_path(name)
    if p is None:
        pytest.skip("fixture not staged")
    q, y0, y1000, y4095, total = KNOWN[name]
    y = O.RefModel(p, quality=q).process_blocks(_sine(), 128)
    tol = 2e-5 if "LSTM" in name or "tw40" in name else 2e-6   # ISA level (AVX2 vs AVX-512) moves the last bits
    assert abs(y[0] - y0) < tol and abs(y[1000] - y1000) < tol and abs(y[4095] - y4095) < tol
    assert abs(float(y.astype(np.float64).sum()) - total) < 4096 * tol
    yp = O.PortModel.from_file(p, quality=q).process(_sine())
    assert float(np.abs(yp - y).max()) <= (5e-5 if tol > 1e-5 else 1e-5)


@pytest.mark.skipif(not O.ref_available(), reason="compiled reference not staged")
@pytest.mark.parametrize("q", [1.0, 0.0])
def test_compiled_reference_a2_container(q):
    p = O.model_path("BossWN-a2.nam")
    if p is None:
        pytest.skip("fixture not staged")
    y0, y1000, y4095, total = KNOWN_A2[q]
    r = O.RefModel(p, quality=q)
    y = r.process_blocks(_sine(), 128)
    assert abs(y[0] - y0) < 2e-6 and abs(y[1000] - y1000) < 2e-6 and abs(y[4095] - y4095) < 2e-6
    assert r.has_quality() and r.receptive_field() == 6346
    yp = O.PortModel.from_file(p, quality=q).process(_sine())
    assert float(np.abs(yp - y).max()) <= 1e-5


@pytest.mark.skipif(not O.ref_available(), reason="compiled reference not staged")
def test_reference_is_chunk_invariant_and_port_agrees_on_fresh_noise():
    p = O.model_path("BossWN-feather.nam")
    if p is None:
        pytest.skip("fixture not staged")
    x = np.random.default_rng(99).uniform(-1, 1, 3000).astype(np.float32)
    a = O.RefModel(p).process_blocks(x, 128)
    b = O.RefModel(p).process_blocks(x, 37)
    assert np.array_equal(a, b)   # bit-identical for any chunking
    c = O.PortModel.from_file(p).process(x)
    assert float(np.abs(a - c).max()) <= 1e-5


def test_port_rejects_wrong_weight_count(tmp_path):
    g = load_golden(golden_files("syn_a1_nano")[0])
    d = dict(g["model"])
    d["weights"] = [float(x) for x in g["weights"]][:-3]
    with pytest.raises(RuntimeError):
        O.PortModel(d)


def test_oversample_rewrites_dilations():
    g = load_golden(golden_files("syn_a1_nano")[0])
    d = json.loads(json.dumps(g["model"]))
    d["weights"] = []
    O.oversample_nam_config(d, 96000)   # NeuralModel.cpp:92-130
    assert d["config"]["layers"][0]["dilations"] == [2, 4, 8, 16, 32, 64, 128]
