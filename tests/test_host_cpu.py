"""CPU: the C-ABI library loads, exports every declared symbol, and the host-side logic (JSON ingest, architecture
dispatch, weight packing, error paths) behaves -- without any compute call (there is no GPU here)."""
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_files, golden_id, load_golden, model_file_for, external_sample_rate_of


def _declared_symbols():
    names = []
    with open(os.path.join(ROOT, "include", "NeuralAudioCApi.h")) as f:
        for line in f:
            m = re.match(r"\s*NA_EXTERN\s+[\w\s\*]+?\b(\w+)\s*\(", line)
            if m:
                names.append(m.group(1))
    return names


REFERENCE_EXPORTS = ["CreateLoader", "DeleteLoader", "CreateModelFromFile", "DeleteModel", "SetLSTMLoadMode", "SetWaveNetLoadMode",
                     "SetAudioInputLevelDBu", "SetDefaultMaxAudioBufferSize", "GetLoadMode", "IsStatic", "SetMaxAudioBufferSize",
                     "GetRecommendedInputDBAdjustment", "GetRecommendedOutputDBAdjustment", "GetSampleRate", "Process"]


def test_library_exports_every_declared_symbol(na):
    L = na.load_library()
    declared = _declared_symbols()
    assert set(REFERENCE_EXPORTS) <= set(declared)      # the reference's 15 exports, NeuralAudioCApi.h:18-46
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), "missing export " + name
    # the Python mirror binds exactly the declared set
    assert set(L._na_signatures) == set(declared)


def test_only_the_public_surface_is_exported(na):
    """Exports = the C ABI of include/NeuralAudioCApi.h plus the four out-of-line NeuralModelLoader::Create* methods of
    include/NeuralAudio/NeuralModel.h (the C++ surface a ModelTest-style caller links against); nothing else leaks."""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "-C", "--defined-only", na.library_path()], text=True)
    syms = [l.split(" T ", 1)[1].strip() for l in out.splitlines() if " T " in l]
    c_syms = [s for s in syms if "::" not in s]
    cpp_syms = sorted(s.split("(")[0] for s in syms if "::" in s)
    assert sorted(c_syms) == sorted(_declared_symbols())
    assert cpp_syms == ["NeuralAudio::b200::NeuralModelLoader::" + m for m in ("CreateFromFile", "CreateFromJsonText", "CreateFromStream", "CreateShardedFromFile")]


def test_loader_handles_work_without_gpu(na):
    L = na.load_library()
    h = L.CreateLoader()
    assert h
    L.SetLSTMLoadMode(h, 1)       # RTNeural: refused silently, like an unsupported mode in the reference
    L.SetWaveNetLoadMode(h, 2)
    L.SetAudioInputLevelDBu(h, 6.0)
    L.SetDefaultMaxAudioBufferSize(h, 256)
    L.DeleteLoader(h)


@pytest.mark.parametrize("path", golden_files(), ids=golden_id)
def test_describe_matches_reference_dispatch(na, path, tmp_path):
    g = load_golden(path)
    mf = model_file_for(g, tmp_path)
    if mf is None:
        pytest.skip("fixture model not staged")
    d = na.describe_model_file(mf, external_sample_rate_of(g))
    info = g["info"]
    if d["kind"] == "container":
        assert info["has_quality"]
        subs = d["submodels"]
        assert [s["max_value"] for s in subs] == [0.5, 1.0]
        assert [s["model"]["arrays"][0]["channels"] for s in subs] == [3, 8]
        assert all(s["model"]["receptive_field"] == 6346 and s["model"]["static"] for s in subs)
        return
    assert d["static"] == info["static"]                   # IsStatic() of the reference for the same file
    if d["kind"] == "wavenet":
        if info["static"]:
            assert d["receptive_field"] == info["rf"]      # 4092 (A1) / 6346 (A2)
        assert d["num_layers"] == sum(a["layers"] for a in d["arrays"])
        assert d["state_floats"] % 4 == 0 and all(lp % 4 == 0 for lp in d["ring_lp"])
    else:
        assert info["rf"] == -1
        assert d["lanes"] >= d["hidden"] and (d["lanes"] in (4, 8, 16, 32) or (d["hidden"] > 32 and d["lanes"] % 4 == 0))


def test_state_size_close_to_algorithmic_minimum(na, tmp_path):
    # A1 Standard keeps 49104 floats of history per stream; ring padding may add < 0.2 %
    g = load_golden(golden_files("syn_a1_standard")[0])
    d = na.describe_model_file(model_file_for(g, tmp_path))
    assert 49104 <= d["state_floats"] <= 49104 * 1.002
    assert d["num_weights"] == 13802


def test_wrong_weight_count_is_reported(na, tmp_path):
    g = load_golden(golden_files("syn_a1_feather")[0])
    d = dict(g["model"])
    d["weights"] = [float(x) for x in g["weights"]][:-1]
    p = tmp_path / "short.nam"
    p.write_text(json.dumps(d))
    with pytest.raises(na.NeuralAudioError, match="Wrong number of weights. Expected 3026 but got 3025"):   # WaveNet.h:704-709 wording
        na.describe_model_file(str(p))


def test_unsupported_models_fail_loudly(na, tmp_path):
    g = load_golden(golden_files("syn_a1_nano")[0])
    d = json.loads(json.dumps(g["model"]))
    d["weights"] = [float(x) for x in g["weights"]]
    d["config"]["layers"][0]["gated"] = True
    p = tmp_path / "gated.nam"
    p.write_text(json.dumps(d))
    with pytest.raises(na.NeuralAudioError, match="gated"):
        na.describe_model_file(str(p))
    p2 = tmp_path / "gru.json"
    p2.write_text(json.dumps({"layers": [{"type": "gru", "shape": [None, None, 8], "weights": []}, {"type": "dense", "shape": [None, None, 1], "weights": []}]}))
    with pytest.raises(na.NeuralAudioError, match="no CPU fallback"):
        na.describe_model_file(str(p2))
    p3 = tmp_path / "bad.nam"
    p3.write_text("{ not json")
    with pytest.raises(na.NeuralAudioError, match="json"):
        na.describe_model_file(str(p3))


def test_a2_with_other_delays_loads_and_other_nam_core_features_do_not(na, tmp_path):
    """SURVEY 8(f4), the part built: an A2 file whose only non-standard property is its timing (what OversampleNAMConfig makes of
    an A2 model on a 96 kHz host: dilations and head dilation doubled, NeuralModel.cpp:92-130) is NAM Core's in the reference and
    loads here with NAM Core's receptive field; every other NAM-Core-only feature is still refused loudly."""
    g = load_golden(golden_files("syn_a2_full_sr96000")[0])
    mf = model_file_for(g, tmp_path)
    d48 = na.describe_model_file(mf, 48000)
    assert d48["static"] is True and d48["receptive_field"] == 6346 and d48["kernel"] == "tcgen05_fp16_pairs"
    d96 = na.describe_model_file(mf, 96000)
    assert d96["static"] is False and d96["receptive_field"] == 2 * 6346 and d96["ring_lp"][-1] == 32     # head history: 15 x 2 frames
    assert g["info"]["rf"] == 2 * 6346 + 1      # NAMModel::GetReceptiveFieldSize of the reference for the same load
    base = json.loads(json.dumps(g["model"]))
    base["weights"] = [float(x) for x in g["weights"]]
    for key, val, why in (("head1x1", {"active": True, "out_channels": 1, "groups": 1}, "non-standard"),
                          ("gating_mode", ["gated"] * 23, "non-standard"),
                          ("conv_pre_film", {"active": True, "shift": True, "groups": 1}, "non-standard"),
                          ("bottleneck", 4, "non-standard"), ("groups_input", 2, "non-standard")):
        d = json.loads(json.dumps(base))
        d["config"]["layers"][0][key] = val
        p = tmp_path / ("a2_%s.nam" % key)
        p.write_text(json.dumps(d))
        with pytest.raises(na.NeuralAudioError, match=why):
            na.describe_model_file(str(p))


def test_oversampling_leaves_the_static_path(na, tmp_path):
    # OversampleNAMConfig (NeuralModel.cpp:92-130): at 96 kHz the dilations double, the file stops being an official shape
    g = load_golden(golden_files("syn_a1_nano")[0])
    mf = model_file_for(g, tmp_path)
    d = na.describe_model_file(mf, 96000)
    assert d["static"] is False and d["receptive_field"] == 2 * 4092


def test_no_gpu_means_loud_failure_not_fallback(na, tmp_path):
    if na.device_count() > 0:
        pytest.skip("a GPU is present")
    g = load_golden(golden_files("syn_a1_nano")[0])
    mf = model_file_for(g, tmp_path)
    with pytest.raises(na.NeuralAudioError, match="no CPU fallback"):
        na.NeuralModelLoader().CreateFromFile(mf)
    L = na.load_library()
    h = L.CreateLoader()
    assert not L.CreateModelFromFile(h, mf)     # NULL, never a half-built model
    assert not L.CreateModelFromFile(h, str(tmp_path / "nope.nam"))
    L.DeleteLoader(h)


def test_product_never_touches_the_oracle():
    # the product path must not import, load or link anything under oracle/
    pkg = os.path.join(ROOT, "neuralaudio_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "na_oracle" not in txt and "libna_ref" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn


def test_tmem_operand_packing_reconstructs_weights(na, tmp_path):
    """Host logic of the TMEM-operand (tcgen05) WaveNet kernel, no GPU: A1 Standard / Lite shapes select it, its packing keeps
    the reference's state size (WaveNet.h:30-83: (K-1)*d columns per conv) with channels padded to (16, 8), and every conv
    tap matrix splits as hi + lo with hi exactly representable in tf32 and |hi + lo - w| at fp32 rounding level."""
    for name, state_floats in (("syn_a1_standard", 49104), ("syn_a1_lite", None)):
        g = load_golden(golden_files(name)[0])
        mf = model_file_for(g, tmp_path)
        d = na.describe_model_file(mf)
        assert d["kernel"] == "tcgen05_fp16_pairs"     # the default; use_tc = 2 selects the 3xTF32 kernel described by d["ts"]
        ts = d["ts"]
        assert ts["num_rings"] == d["num_rings"] == d["num_layers"]
        pad = [a["padded"] for a in d["arrays"]]
        cfg = g["model"]["config"]["layers"]
        assert ts["state_floats"] == sum(cpad * (a["kernel_size"] - 1) * dil for cpad, a in zip((16, 8), cfg) for dil in a["dilations"])
        if state_floats:
            assert ts["state_floats"] == state_floats and pad[0] >= 16
        assert ts["conv_hi_not_tf32"] == 0
        # fp16-pair packing: same state size (the rings hold the packed (h1, h2) pairs, 4 bytes per value), every conv tap
        # reconstructs as W1 + W2 to half an fp16 subnormal step (2^-25) or 2^-22 relative, whichever is larger
        h = d["h"]
        assert h["state_floats"] == ts["state_floats"] and h["num_rings"] == ts["num_rings"]
        assert h["conv_split_max_error"] <= 2.0 ** -22 * max(1.0, float(np.abs(g["weights"]).max()))
        # four streams per SM (window regions of consecutive layers are disjoint, so the buffer is 512 rows): shared memory per
        # CTA must stay below (228 KB - 4 KB) / 4
        assert h["win_rows"] * 64 + 2 * h["max_block_bytes"] + h["table_bytes"] + 2048 < (228 * 1024 - 4 * 1024) // 4
        assert ts["conv_split_max_error"] <= 1e-7
        assert ts["max_block"] * 4 * 2 <= 48 * 1024      # two weight buffers per CTA, 4 CTAs per SM
    g = load_golden(golden_files("syn_a1_nano")[0])
    assert na.describe_model_file(model_file_for(g, tmp_path))["kernel"] == "cuda_cores"


def test_cpp_consumer_builds_links_and_fails_loudly_without_gpu(tmp_path):
    """A C++ caller of the reference (ModelTest.cpp:11-57 style) recompiles against include/NeuralAudio/NeuralModel.h and
    links the shared library: the loader's out-of-line methods are exported.  Without a CUDA device loading returns nullptr
    (no CPU fallback), which the tool reports the way the reference's ModelTest does."""
    import subprocess
    import __graft_entry__ as ge
    exe = ge.build_model_test()
    r = subprocess.run([exe, str(tmp_path / "missing.nam")], capture_output=True, text=True, timeout=120)
    assert "Model file does not exist" in r.stdout and r.returncode == 1
    import torch
    if not torch.cuda.is_available():
        g = load_golden(golden_files("syn_a1_nano")[0])
        mf = model_file_for(g, tmp_path)
        r = subprocess.run([exe, mf], capture_output=True, text=True, timeout=120)
        assert "Unable to load model from" in r.stdout and r.returncode == 1


def test_lstm_kernel_choice_by_shape(na, tmp_path):
    """Host logic of the LSTM dispatch (reported for a model of 8192 stream slots, and of 32768): gate rows in registers where they
    fit (up to 16 units in one layer, 8 in two; 1x16 moves to the tensor-core kernel once the batch is large enough for that kernel's
    fixed step chain to pay, ~14000 streams); past that register cliff the tensor-core kernel (one or two layers, up to 32 units) from ~5000 streams, the lane-per-stream
    kernel with shared-memory matrices for everything else (three or more layers, more than 32 units)."""
    want = {"syn_lstm_1x16": ("lstm_gate_rows_in_registers", "lstm_tcgen05_gates"), "syn_lstm_2x8": ("lstm_gate_rows_in_registers", "lstm_gate_rows_in_registers"),
            "syn_lstm_1x8": ("lstm_gate_rows_in_registers", "lstm_gate_rows_in_registers"),
            "syn_lstm_1x24": ("lstm_tcgen05_gates", "lstm_tcgen05_gates"), "syn_lstm_2x12": ("lstm_tcgen05_gates", "lstm_tcgen05_gates"),
            "syn_dyn_lstm_2x32": ("lstm_tcgen05_gates", "lstm_tcgen05_gates"),
            "syn_dyn_lstm_3x18": ("lstm_lane_per_stream",) * 2, "syn_dyn_lstm_1x40": ("lstm_lane_per_stream",) * 2, "syn_dyn_lstm_4x6": ("lstm_lane_per_stream",) * 2}
    for name, (kernel, kernel_large) in want.items():
        g = load_golden(golden_files(name)[0])
        d = na.describe_model_file(model_file_for(g, tmp_path))
        assert d["kernel"] == kernel, (name, d["kernel"])
        assert d["kernel_32768_streams"] == kernel_large, (name, d["kernel_32768_streams"])


def test_lstm_tensor_core_kernel_refuses_weights_outside_fp16_range(na, tmp_path):
    """The tensor-core LSTM kernel holds the gate matrices as fp16 pairs: a model with a weight beyond the fp16 range keeps the
    fp32 CUDA-core kernels at every batch size (PackLstm's tcOk)."""
    import json
    g = load_golden(golden_files("syn_lstm_1x24")[0])
    w = np.asarray(g["weights"], dtype=np.float32).copy()
    w[5] = 1.0e6
    d = dict(g["model"]); d["weights"] = [float(v) for v in w]
    mf = os.path.join(str(tmp_path), "huge_weight.nam")
    with open(mf, "w") as f:
        json.dump(d, f)
    assert na.describe_model_file(mf)["kernel_32768_streams"] == "lstm_lane_per_stream"
