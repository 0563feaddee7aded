"""GPU: parity of the CUDA path against the oracle, called through the C ABI (neuralaudio_b200 mirrors the binding).

Checker = committed golden vectors (outputs of the unmodified reference) + the plain-C oracle on fresh seeded inputs.
Tolerances are stated in conftest.py (WaveNet 1e-5 max-abs per BASELINE.json north_star; LSTM 5e-5)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import WAVENET_TOL, LSTM_TOL, golden_files, golden_id, load_golden, model_file_for, tol_for, is_lstm_case, external_sample_rate_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _load(na, mf, quality=1.0, streams=1, prewarm=True, sample_rate=48000):
    ld = na.NeuralModelLoader()
    ld.SetExternalSampleRate(sample_rate)
    ld.SetDefaultQualityScaleFactor(quality)
    ld.SetDefaultNumStreams(streams)
    return ld.CreateFromFile(mf, prewarm)


def _blocks(model, x, block):
    y = np.empty_like(x)
    for i in range(0, x.size, block):
        y[i:i + block] = model.Process(np.ascontiguousarray(x[i:i + block]))
    return y


@pytest.mark.parametrize("path", golden_files(), ids=golden_id)
def test_single_stream_process_matches_reference_golden(na, path, tmp_path):
    """cfg 1 (and every other shape): the reference's own Process() call sequence, 128-frame host buffers."""
    g = load_golden(path)
    mf = model_file_for(g, tmp_path)
    if mf is None:
        pytest.skip("fixture model not staged")
    q = float(g.get("quality", 1.0))
    sr = external_sample_rate_of(g)
    m = _load(na, mf, q, sample_rate=sr)
    y = _blocks(m, g["x"], 128)
    err = float(np.abs(y - g["y"]).max())
    assert err <= tol_for(g), "max-abs vs reference %.3g" % err
    # boundary metadata equals the reference's answers for the same file
    info = g["info"]
    assert m.IsStatic() == info["static"]
    assert m.GetReceptiveFieldSize() == info["rf"]
    assert m.GetLoadMode() == 0
    assert abs(m.GetSampleRate() - info["sample_rate"]) < 1e-3
    assert abs(m.GetRecommendedInputDBAdjustment() - info["in_adj"]) < 1e-4
    assert abs(m.GetRecommendedOutputDBAdjustment() - info["out_adj"]) < 1e-4
    assert m.HasQualityScaling() == info["has_quality"]
    # zero input after load: the prewarmed steady state
    m2 = _load(na, mf, q, sample_rate=sr)
    dc = _blocks(m2, np.zeros(512, dtype=np.float32), 128)
    assert float(np.abs(dc - g["dc"]).max()) <= tol_for(g)


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_a1_nano", "syn_a2_full", "syn_lstm_1x16", "syn_lstm_2x8"])
def test_chunking_invariance_and_inplace(na, name, tmp_path):
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    x = g["x"][:3000]
    ref = _blocks(_load(na, mf), x, 128)
    assert float(np.abs(ref - g["y"][:3000]).max()) <= tol_for(g)
    for block in (1, 37, 64, 200, 3000):
        if block == 1:
            y = _blocks(_load(na, mf), x[:300], 1)
            assert np.array_equal(y, ref[:300]), "block 1"
            continue
        y = _blocks(_load(na, mf), x, block)
        assert np.array_equal(y, ref), "block %d differs by %.3g" % (block, np.abs(y - ref).max())
    # n = 0 is a no-op; in == out is allowed (WaveNet.h:770, LSTM.h:168,184)
    m = _load(na, mf)
    m.Process(np.zeros(0, dtype=np.float32))
    buf = x.copy()
    for i in range(0, buf.size, 128):
        seg = buf[i:i + 128]
        m.Process(seg, seg)
    assert np.array_equal(buf, ref)


@pytest.mark.parametrize("name,streams,calls,frames", [("syn_a1_standard", 40, 12, 128), ("syn_a1_feather", 19, 10, 128),
                                                        ("syn_a2_full", 24, 6, 256), ("syn_a2_lite", 9, 6, 256),
                                                        ("syn_lstm_1x16", 70, 8, 128), ("syn_lstm_2x12", 21, 6, 128),
                                                        ("syn_a1_lite", 11, 8, 96)])
def test_batch_matches_oracle_per_stream(na, O, name, streams, calls, frames, tmp_path):
    """Many independent streams per launch, state carried across calls, ragged stream counts; every stream is checked
    against its own oracle instance on the same seeded white noise."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    rng = np.random.default_rng(4242)
    amp = 0.5 if is_lstm_case(g) else 1.0
    x = (rng.uniform(-1, 1, (calls, streams, frames)) * amp).astype(np.float32)
    m = _load(na, mf, streams=streams)
    assert m.GetNumStreams() == streams
    y = np.empty_like(x)
    for c in range(calls):
        m.ProcessBatch(x[c], y[c], streams, frames, na.STREAM_MAJOR)
    worst = 0.0
    for s in range(streams):
        om = O.PortModel.from_file(mf)
        ys = om.process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        worst = max(worst, float(np.abs(ys - y[:, s, :].reshape(-1)).max()))
    assert worst <= tol_for(g), "worst stream max-abs %.3g" % worst


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_lstm_1x16", "syn_a2_full"])
def test_layouts_and_device_pointers_agree(na, name, tmp_path):
    import torch
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 33, 128, 4
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (calls, S, n)).astype(np.float32)
    a = _load(na, mf, streams=S)
    b = _load(na, mf, streams=S)
    c = _load(na, mf, streams=S)
    ya = np.empty_like(x)
    yb = np.empty((calls, n, S), dtype=np.float32)
    yc = torch.empty((calls, S, n), dtype=torch.float32, device="cuda")
    xc = torch.from_numpy(x).cuda()
    for k in range(calls):
        a.ProcessBatch(x[k], ya[k], S, n, na.STREAM_MAJOR)
        b.ProcessBatch(np.ascontiguousarray(x[k].T), yb[k], S, n, na.FRAME_MAJOR)
        c.ProcessBatch(xc[k], yc[k], S, n, na.STREAM_MAJOR)   # device pointers, asynchronous
    c.Synchronize()
    assert np.array_equal(ya, np.transpose(yb, (0, 2, 1)))
    assert np.array_equal(ya, yc.cpu().numpy())


def test_prewarm_semantics(na, tmp_path):
    # WaveNet: Prewarm() again == full reset; no-prewarm load starts from zero history
    g = load_golden(golden_files("syn_a1_feather")[0])
    mf = model_file_for(g, tmp_path)
    m = _load(na, mf)
    y1 = _blocks(m, g["x"][:1024], 128)
    m.Prewarm()
    y2 = _blocks(m, g["x"][:1024], 128)
    assert np.array_equal(y1, y2)
    cold = _blocks(_load(na, mf, prewarm=False), g["x"][:1024], 128)
    assert float(np.abs(cold - y1).max()) > 1e-3
    # LSTM: Prewarm() again == 2048 more zero samples from the current state, not a reset
    g = load_golden(golden_files("syn_lstm_1x16")[0])
    mf = model_file_for(g, tmp_path)
    m = _load(na, mf)
    a = _blocks(m, g["x"][:512], 128)
    m.Prewarm()
    b = _blocks(m, g["x"][:512], 128)
    ref = _load(na, mf)
    a2 = _blocks(ref, g["x"][:512], 128)
    _blocks(ref, np.zeros(2048, dtype=np.float32), 64)
    b2 = _blocks(ref, g["x"][:512], 128)
    assert np.array_equal(a, a2)
    assert float(np.abs(b - b2).max()) <= 1e-6


def test_a2_container_quality_switch(na, O, tmp_path):
    p = O.model_path("BossWN-a2.nam")
    if p is None:
        pytest.skip("fixture not staged")
    full = load_golden(golden_files("ref_BossWN_a2.")[0])
    lite = load_golden(golden_files("ref_BossWN_a2_q0")[0])
    m = _load(na, p, 1.0)
    assert m.HasQualityScaling() and m.GetReceptiveFieldSize() == 6346
    y = _blocks(m, full["x"][:2048], 256)
    assert float(np.abs(y - full["y"][:2048]).max()) <= WAVENET_TOL
    # quality 0.0..0.5 selects the 3-channel sub-model, whose state is still the prewarmed one (LoadAll)
    m.SetQualityScaleFactor(0.3)
    assert abs(m.GetQualityScaleFactor() - 0.3) < 1e-7
    y = _blocks(m, lite["x"][:2048], 256)
    assert float(np.abs(y - lite["y"][:2048]).max()) <= WAVENET_TOL
    # container-level metadata (CompositeModel.h:130-135)
    assert abs(m.GetRecommendedOutputDBAdjustment() - full["info"]["out_adj"]) < 1e-4
    assert m.GetMetadata("gear_type") == '"pedal"' and m.GetMetadata("nope") == ""
    assert m.GetModelVersion() == "0.7.0"


def test_metadata_matches_reference(na, O, tmp_path):
    p = O.model_path("BossWN-standard.nam")
    if p is None or not O.ref_available():
        pytest.skip("fixture / compiled reference not staged")
    r = O.RefModel(p)
    m = _load(na, p)
    for key in ("name", "loudness", "gain", "date", "modeled_by", "input_level_dbu", "training", "missing"):
        assert m.GetMetadata(key) == r.metadata(key), key
    assert m.GetModelVersion() == r.version()


def test_tma_and_plain_window_paths_are_bit_identical(na, tmp_path):
    g = load_golden(golden_files("syn_a1_standard")[0])
    mf = model_file_for(g, tmp_path)
    S, n = 16, 128
    x = np.random.default_rng(3).uniform(-1, 1, (6, S, n)).astype(np.float32)
    outs = []
    for tma in (1, 0):
        prev = na.set_option("use_tma", tma)
        try:
            m = _load(na, mf, streams=S)
            y = np.empty_like(x)
            for k in range(x.shape[0]):
                m.ProcessBatch(x[k], y[k], S, n)
            outs.append(y)
        finally:
            na.set_option("use_tma", prev)
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("name", ["syn_a1_nano.", "syn_a2_lite", "ref_BossWN_nano", "syn_dyn_single6_k2"])
def test_single_stream_kernel_is_bit_identical_to_batched_kernel(na, name, tmp_path):
    """A single-stream call of a small WaveNet runs on the one-CTA kernel (whole ring state in shared memory, wavenet_one_kernels.cu):
    its results equal the batched kernel's bit for bit - for buffers of 1, 37 and 128 frames and for calls longer than one pass -
    and a stream advanced partly by Process() and partly as slot 0 of ProcessBatch() stays consistent."""
    files = golden_files(name)
    if not files:
        pytest.skip("no such vector")
    g = load_golden(files[0])
    mf = model_file_for(g, tmp_path)
    if mf is None:
        pytest.skip("fixture model not staged")
    q = float(g.get("quality", 1.0))
    x = np.random.default_rng(11).uniform(-1, 1, 128 * 7 + 37 + 1 + 300).astype(np.float32)
    cuts = [0, 128, 256, 293, 294, 422, 722, 850, 978, 1106, x.size]
    outs = {}
    for one in (1, 0):
        prev = na.set_option("use_one", one)
        try:
            m = _load(na, mf, q)
            y = np.empty_like(x)
            for a, b in zip(cuts[:-1], cuts[1:]):
                y[a:b] = m.Process(np.ascontiguousarray(x[a:b]))
            outs[one] = y
        finally:
            na.set_option("use_one", prev)
    assert np.array_equal(outs[1], outs[0])
    # mixed use: batch of 3 slots, slot 0 advanced alternately through Process() and ProcessBatch()
    m = _load(na, mf, q, streams=3)
    xb = np.random.default_rng(12).uniform(-1, 1, (4, 3, 128)).astype(np.float32)
    ref = _load(na, mf, q, streams=3)
    for c in range(4):
        yb = np.empty_like(xb[c]); ref.ProcessBatch(xb[c], yb, 3, 128)
        if c % 2 == 0:
            y0 = m.Process(np.ascontiguousarray(xb[c, 0]))      # slot 0 only (slots 1, 2 of `m` fall behind: not compared)
        else:
            yy = np.empty_like(xb[c]); m.ProcessBatch(xb[c], yy, 3, 128); y0 = yy[0]
        assert np.array_equal(y0, yb[0])


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_lstm_1x16"])
def test_pipelined_host_path_matches_blocking(na, name, tmp_path):
    """NA_ProcessBatchAsync (copy-in / kernels / copy-out of consecutive calls overlapped) is the same computation as the
    blocking host path: bit-identical outputs, and pageable buffers are refused loudly."""
    import torch
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 24, 128, 7
    x = np.random.default_rng(23).uniform(-1, 1, (calls, S, n)).astype(np.float32)
    m = _load(na, mf, streams=S)
    yb = np.empty_like(x)
    for k in range(calls):
        m.ProcessBatch(x[k], yb[k], S, n)
    m2 = _load(na, mf, streams=S)
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.empty_like(xp).pin_memory()
    for k in range(calls):
        m2.ProcessBatchAsync(xp[k], yp[k], S, n)
        if k >= 1:
            m2.WaitBatches(1)
            assert np.array_equal(yp[k - 1].numpy(), yb[k - 1])
    m2.WaitBatches(0)
    assert np.array_equal(yp.numpy(), yb)
    with pytest.raises(na.NeuralAudioError, match="page-locked"):
        m2.ProcessBatchAsync(x[0], np.empty_like(x[0]), S, n)


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_a1_lite"])
def test_tensor_core_and_cuda_core_kernels_agree(na, O, name, tmp_path):
    """The tcgen05 (3xTF32) kernel and the CUDA-core fp32 kernel are two independent implementations of the same
    path with different state layouts; both must sit within tolerance of the oracle and of each other."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 12, 128, 10
    x = np.random.default_rng(17).uniform(-1, 1, (calls, S, n)).astype(np.float32)
    outs = []
    for tc in (2, 0, 3):   # 3xTF32 TMEM-operand tcgen05 kernel, CUDA-core kernel, fp16-pair tcgen05 kernel (default)
        prev = na.set_option("use_tc", tc)
        try:
            m = _load(na, mf, streams=S)
            y = np.empty_like(x)
            for k in range(calls):
                m.ProcessBatch(x[k], y[k], S, n)
            outs.append(y)
        finally:
            na.set_option("use_tc", prev)
    assert float(np.abs(outs[0] - outs[1]).max()) <= 4e-6
    assert float(np.abs(outs[2] - outs[1]).max()) <= 4e-6
    for s in (0, S - 1):
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        for y in outs:
            assert float(np.abs(ys - y[:, s, :].reshape(-1)).max()) <= WAVENET_TOL


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_a1_feather", "syn_a2_full", "syn_dyn_20x10", "syn_dyn_16x16_k5",
                                  "syn_dyn_3arrays", "syn_dyn_4arrays", "syn_dyn_48x24", "syn_dyn_single40_k3"])
def test_runtime_shaped_kernel(na, O, name, tmp_path):
    """The run-time-shaped kernels (the batched counterpart of the reference's dynamic path, WaveNetDynamic.h) against the
    oracle per stream, for shapes that only they can run (three and four layer arrays; more than 32 channels: the
    shared-memory form) and - forced with use_tc = -1 - for official shapes, where they must also agree with the
    specialised kernels."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 9, 100, 11     # ragged: 100-frame calls
    x = np.random.default_rng(29).uniform(-1, 1, (calls, S, n)).astype(np.float32)
    outs = []
    for tc in (-1, 2):
        prev = na.set_option("use_tc", tc)
        try:
            m = _load(na, mf, streams=S)
            y = np.empty_like(x)
            for k in range(calls):
                m.ProcessBatch(x[k], y[k], S, n)
            outs.append(y)
        finally:
            na.set_option("use_tc", prev)
    assert float(np.abs(outs[0] - outs[1]).max()) <= 4e-6
    for s in (0, S // 2, S - 1):
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        assert float(np.abs(ys - outs[0][:, s, :].reshape(-1)).max()) <= WAVENET_TOL


def test_split_launch_matches_fused(na, O, tmp_path):
    """The TS kernel's split form (one launch per layer array, hand-over through a scratch buffer, head sum in registers)
    computes the same path as the fused kernel."""
    g = load_golden(golden_files("syn_a1_standard")[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 20, 128, 9
    x = np.random.default_rng(31).uniform(-1, 1, (calls, S, n)).astype(np.float32)
    outs = []
    prev_tc = na.set_option("use_tc", 2)       # the 3xTF32 kernel (the default is the fp16-pair kernel, which has no split form)
    try:
        for split in (1, 0):
            prev = na.set_option("ts_split", split)
            try:
                m = _load(na, mf, streams=S)
                y = np.empty_like(x)
                for k in range(calls):
                    m.ProcessBatch(x[k], y[k], S, n)
                outs.append(y)
            finally:
                na.set_option("ts_split", prev)
    finally:
        na.set_option("use_tc", prev_tc)
    assert float(np.abs(outs[0] - outs[1]).max()) <= 2e-6
    ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, 3, :]).reshape(-1))
    assert float(np.abs(ys - outs[0][:, 3, :].reshape(-1)).max()) <= WAVENET_TOL


@pytest.mark.parametrize("name,streams", [("syn_lstm_1x16", 37), ("syn_lstm_2x12", 10), ("syn_dyn_lstm_3x18", 23), ("syn_dyn_lstm_1x40", 7),
                                          ("syn_dyn_lstm_2x32", 9)])
def test_runtime_shaped_lstm_kernel(na, O, name, streams, tmp_path):
    """The run-time-shaped LSTM kernel (the reference's dynamic path, LSTMDynamic.h): sizes outside the compile-time-shaped
    list run on it by themselves; use_tc = -1 forces it for the others.  Ragged stream counts (not a multiple of the streams
    one block carries), two layouts, every stream against its own oracle instance."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    calls, n = 5, 96
    x = (np.random.default_rng(41).uniform(-1, 1, (calls, streams, n)) * 0.5).astype(np.float32)
    prev = na.set_option("use_tc", -1)
    try:
        m = _load(na, mf, streams=streams)
        y = np.empty_like(x)
        for k in range(calls):
            m.ProcessBatch(x[k], y[k], streams, n)
        m2 = _load(na, mf, streams=streams)
        yt = np.empty((calls, n, streams), dtype=np.float32)
        for k in range(calls):
            m2.ProcessBatch(np.ascontiguousarray(x[k].T), yt[k], streams, n, na.FRAME_MAJOR)
    finally:
        na.set_option("use_tc", prev)
    assert np.array_equal(y, yt.transpose(0, 2, 1))
    for s in sorted({0, streams // 2, streams - 1}):
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        assert float(np.abs(ys - y[:, s, :].reshape(-1)).max()) <= LSTM_TOL
    if "dyn" not in name:
        # same shapes on the compile-time-shaped kernel
        m3 = _load(na, mf, streams=streams)
        y3 = np.empty_like(x)
        for k in range(calls):
            m3.ProcessBatch(x[k], y3[k], streams, n)
        assert float(np.abs(y3 - y).max()) <= LSTM_TOL


@pytest.mark.parametrize("name,streams", [("syn_lstm_1x16", 70), ("syn_lstm_1x24", 65), ("syn_lstm_2x12", 37), ("syn_lstm_2x8", 129),
                                          ("syn_dyn_lstm_3x18", 23), ("syn_dyn_lstm_1x40", 7), ("syn_dyn_lstm_2x32", 64), ("syn_dyn_lstm_4x6", 3)])
def test_lane_per_stream_lstm_kernel(na, O, name, streams, tmp_path):
    """The lane = stream LSTM kernel (gate matrices once per CTA in shared memory; the default past the register cliff: 1x24,
    2x12, 2x16, run-time sizes up to 64 units) forced for every shape: ragged stream counts around its 64-stream CTAs, both
    layouts, odd call sizes, every probed stream against its own oracle instance; and against the gate-rows-in-registers kernel."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    sizes = [96, 1, 37, 128, 70]
    rng = np.random.default_rng(43)
    xs = [(rng.uniform(-1, 1, (streams, n)) * 0.5).astype(np.float32) for n in sizes]
    outs = {}
    for kern in (2, 1):
        prev = na.set_option("lstm_kernel", kern)
        try:
            m = _load(na, mf, streams=streams)
            m2 = _load(na, mf, streams=streams)
            ys, yts = [], []
            for x in xs:
                y = np.empty_like(x)
                m.ProcessBatch(x, y, streams, x.shape[1])
                ys.append(y)
                yt = np.empty((x.shape[1], streams), dtype=np.float32)
                m2.ProcessBatch(np.ascontiguousarray(x.T), yt, streams, x.shape[1], na.FRAME_MAJOR)
                yts.append(yt.T)
        finally:
            na.set_option("lstm_kernel", prev)
        outs[kern] = np.concatenate(ys, axis=1)
        assert np.array_equal(outs[kern], np.concatenate(yts, axis=1))
    for s in sorted({0, streams // 2, streams - 1}):
        ref = O.PortModel.from_file(mf).process(np.concatenate([x[s] for x in xs]))
        assert float(np.abs(ref - outs[2][s]).max()) <= LSTM_TOL
    assert float(np.abs(outs[1] - outs[2]).max()) <= LSTM_TOL   # (kernel 1 falls back to the automatic choice where it has no variant)


@pytest.mark.parametrize("name,streams", [("syn_lstm_1x16", 70), ("syn_lstm_1x24", 65), ("syn_lstm_2x12", 37), ("syn_lstm_2x8", 129),
                                          ("syn_lstm_2x16", 200), ("syn_dyn_lstm_2x32", 64), ("syn_lstm_1x8", 3), ("ref_BossLSTM_1x16", 131),
                                          ("syn_lstm_1x16", 300), ("syn_lstm_2x8", 257)])
def test_tensor_core_lstm_kernel(na, O, name, streams, tmp_path):
    """The tcgen05 LSTM kernel (gates of 128 streams as one small GEMM per step and layer, fp16-pair operands; the default
    for large batches) forced for every shape it covers, with one and with two 128-stream sets per CTA: ragged stream counts around its CTAs,
    both layouts, odd call sizes across its 16-frame tiles, probed streams against their own oracle instances, the whole batch
    against the fp32 CUDA-core kernels, and bit-identical results however the samples are cut into calls."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    if mf is None:
        pytest.skip("fixture not staged")
    sizes = [96, 1, 37, 128, 70, 16, 17]
    rng = np.random.default_rng(47)
    xs = [(rng.uniform(-1, 1, (streams, n)) * 0.5).astype(np.float32) for n in sizes]
    xall = np.concatenate(xs, axis=1)
    outs = {}
    for kern, sets in ((4, 1), (4, 2), (0, 0)):
        prev = na.set_option("lstm_kernel", kern)
        prev_sets = na.set_option("lstm_tc_sets", sets)
        try:
            m = _load(na, mf, streams=streams)
            m2 = _load(na, mf, streams=streams)
            ys, yts = [], []
            for x in xs:
                y = np.empty_like(x)
                m.ProcessBatch(x, y, streams, x.shape[1])
                ys.append(y)
                yt = np.empty((x.shape[1], streams), dtype=np.float32)
                m2.ProcessBatch(np.ascontiguousarray(x.T), yt, streams, x.shape[1], na.FRAME_MAJOR)
                yts.append(yt.T)
            y1 = np.concatenate(ys, axis=1)
            assert np.array_equal(y1, np.concatenate(yts, axis=1))
            if kern == 4:
                m3 = _load(na, mf, streams=streams)
                yall = np.empty_like(xall)
                m3.ProcessBatch(xall, yall, streams, xall.shape[1])
                assert np.array_equal(y1, yall), "the cut of the call sequence changed the result"
        finally:
            na.set_option("lstm_kernel", prev)
            na.set_option("lstm_tc_sets", prev_sets)
        outs[(kern, sets)] = y1
    # one and two 128-stream sets per CTA: the same arithmetic per stream, bit for bit
    assert np.array_equal(outs[(4, 1)], outs[(4, 2)])
    for s in sorted({0, streams // 2, streams - 1}):
        ref = O.PortModel.from_file(mf).process(xall[s])
        assert float(np.abs(ref - outs[(4, 1)][s]).max()) <= LSTM_TOL
    assert float(np.abs(outs[(4, 1)] - outs[(0, 0)]).max()) <= LSTM_TOL


@pytest.mark.parametrize("kind", ["zero", "huge", "tiny"])
def test_lstm_activation_edge_ranges(na, O, kind, tmp_path):
    """The LSTM kernel computes its gate activations in packed pairs with a hand-scheduled IEEE quotient that is valid
    for ordinary arguments and hands zero / denormal / huge arguments to the scalar IEEE division: drive each range
    (all-zero weights: gates exactly 0; weights x 1e7: saturated gates; weights x 1e-30: gates near the denormals)."""
    import json
    g = load_golden(golden_files("syn_lstm_1x16")[0])
    w = np.asarray(g["weights"], dtype=np.float32).copy()
    scale = {"zero": 0.0, "huge": 1e7, "tiny": 1e-30}[kind]
    H = 16
    w[:4 * H * (1 + H) + 4 * H] *= np.float32(scale)   # gate matrices and biases only: initial state and head stay as they are
    d = dict(g["model"]); d["weights"] = [float(v) for v in w]
    mf = os.path.join(str(tmp_path), "edge_%s.nam" % kind)
    with open(mf, "w") as f:
        json.dump(d, f)
    x = np.random.default_rng(5).uniform(-0.5, 0.5, 1024).astype(np.float32)
    y = _blocks(_load(na, mf), x, 128)
    yo = O.PortModel.from_file(mf).process(x)
    assert np.isfinite(yo).all() and np.isfinite(y).all()
    assert float(np.abs(y - yo).max()) <= LSTM_TOL


# known answers of the reference's Internal path (x[i] = sin(i * 0.01), 4096 samples in 128-sample
# calls on a freshly loaded model): out[0], out[1000], out[4095], sum(out)
KNOWN_ANSWERS = {
    ("BossWN-nano.nam", 1.0): (0.000353399199, -0.252702147, -0.283021569, 21.8273232),
    ("BossWN-feather.nam", 1.0): (-0.00016338109, -0.2728616, -0.295489967, 28.6872998),
    ("BossWN-standard.nam", 1.0): (-0.00067000452, -0.317638844, -0.349413633, 39.1764422),
    ("BossWN-a2.nam", 1.0): (0.000276284292, -0.257261902, -0.284163564, 30.4085193),
    ("BossWN-a2.nam", 0.0): (0.000122590631, -0.236403778, -0.282012135, 14.3841253),
    ("BossLSTM-1x16.nam", 1.0): (-0.0270614624, -0.253457189, -0.171944439, -219.212568),
    ("BossLSTM-2x8.nam", 1.0): (0.00714398921, -0.170382544, -0.184256151, -172.00273),
}


@pytest.mark.parametrize("fixture,quality", sorted(KNOWN_ANSWERS), ids=lambda v: str(v))
def test_cpp_model_test_known_answers(fixture, quality):
    """The C++ surface end to end: tools/model_test (ModelTest.cpp's protocol against our header, plain g++) loads the
    reference's own fixture, runs the known-answer input through Process(), a batch through ProcessBatch(), and its
    batch-vs-single RMS (ComputeError protocol, ModelTest.cpp:81-118) is zero: every stream slot is the same machine."""
    import re
    import subprocess
    import __graft_entry__ as ge
    mf = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "models", fixture)
    if not os.path.exists(mf):
        pytest.skip("fixture model not staged")
    exe = ge.build_model_test()
    r = subprocess.run([exe, "--kat", "-s", "48", "-q", str(quality), mf], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"KAT out\[0\]=(\S+) out\[1000\]=(\S+) out\[4095\]=(\S+) sum=(\S+)", r.stdout)
    assert m, r.stdout
    got = [float(v) for v in m.groups()]
    want = KNOWN_ANSWERS[(fixture, quality)]
    tol = LSTM_TOL if "LSTM" in fixture else WAVENET_TOL
    for a, b in zip(got[:3], want[:3]):
        assert abs(a - b) <= tol, (got, want)
    assert abs(got[3] - want[3]) <= 4096 * tol
    assert re.search(r"Internal: \S+ \(\S+xRT\)", r.stdout)
    assert re.search(r"Batch 48 streams: \S+ \(\S+xRT, \S+ Msamples/s\)", r.stdout)
    rms = float(re.search(r"Batch vs single RMS err: (\S+)", r.stdout).group(1))
    assert rms <= 1e-6, r.stdout


def test_full_size_config_properties(na, O, tmp_path):
    """BASELINE.json cfg 2 (A1 Standard, 4096 streams x 128 frames) at full size, through size-independent
    properties: identical inputs => bit-identical streams (independence + determinism), a tile of streams checked
    sample-by-sample against the oracle, and linearity of the stream <-> slot mapping (reversed batch => reversed output)."""
    import torch
    g = load_golden(golden_files("syn_a1_standard")[0])
    mf = model_file_for(g, tmp_path)
    S, n, calls = 4096, 128, 6
    rng = np.random.default_rng(11)
    base = rng.uniform(-1, 1, (calls, 8, n)).astype(np.float32)
    x = np.tile(base, (1, S // 8, 1))           # stream s carries pattern s % 8
    m = _load(na, mf, streams=S)
    assert m.GetStateBytesPerStream() >= 196416
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    for k in range(calls):
        m.ProcessBatch(xd[k], yd[k], S, n)
    m.Synchronize()
    y = yd.cpu().numpy()
    for p in range(8):
        assert np.array_equal(y[:, p::8, :], np.broadcast_to(y[:, p:p + 1, :], y[:, p::8, :].shape))
        om = O.PortModel.from_file(mf)
        ys = om.process(np.ascontiguousarray(base[:, p, :]).reshape(-1))
        assert float(np.abs(ys - y[:, p, :].reshape(-1)).max()) <= WAVENET_TOL
    m2 = _load(na, mf, streams=S)
    xr = torch.flip(xd, dims=[1]).contiguous()
    yr = torch.empty_like(xr)
    for k in range(calls):
        m2.ProcessBatch(xr[k], yr[k], S, n)
    m2.Synchronize()
    assert torch.equal(torch.flip(yr, dims=[1]), yd)


@pytest.mark.parametrize("name,streams", [("syn_a1_standard_sr96000", 20), ("syn_a1_nano_sr96000", 33), ("syn_a2_full_sr96000", 21), ("syn_a2_lite_sr96000", 9)])
def test_oversampled_batch_matches_oracle(na, O, name, streams, tmp_path):
    """Host at 96 kHz: every dilation doubles (OversampleNAMConfig, NeuralModel.cpp:92-130), the receptive field becomes 8184
    and the rings twice as long; the batch kernels run the doubled dilations as run-time values."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    sr = external_sample_rate_of(g)
    calls, n = 70, 128                     # 8960 frames > the 8184-frame receptive field
    x = np.random.default_rng(97).uniform(-1, 1, (calls, streams, n)).astype(np.float32)
    m = _load(na, mf, streams=streams, sample_rate=sr)
    y = np.empty_like(x)
    for k in range(calls):
        m.ProcessBatch(x[k], y[k], streams, n)
    for s in (0, streams - 1):
        ys = O.PortModel.from_file(mf, external_sample_rate=sr).process(np.ascontiguousarray(x[:, s, :]).reshape(-1))
        assert float(np.abs(ys - y[:, s, :].reshape(-1)).max()) <= WAVENET_TOL


@pytest.mark.parametrize("name", ["syn_a1_standard", "syn_a2_full", "syn_lstm_1x16", "syn_dyn_7x3", "syn_dyn_lstm_3x18"])
def test_empty_partial_and_long_calls(na, O, name, tmp_path):
    """Ragged use of the batch API: an empty call changes nothing; a call on the first k of S slots advances only those;
    one very long call (many internal passes) equals the same audio in 128-frame calls."""
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    tol = tol_for(g)
    S, n = 12, 128
    amp = 0.5 if is_lstm_case(g) else 1.0
    x = (np.random.default_rng(77).uniform(-1, 1, (3, S, n)) * amp).astype(np.float32)
    m = _load(na, mf, streams=S)
    y = np.empty_like(x)
    m.ProcessBatch(x[0], y[0], S, n)
    m.ProcessBatch(np.empty((S, 0), dtype=np.float32), np.empty((S, 0), dtype=np.float32), S, 0)       # empty call
    m.Process(np.empty(0, dtype=np.float32))
    k = 5
    m.ProcessBatch(np.ascontiguousarray(x[1][:k]), y[1][:k], k, n)                                     # first k slots only
    m.ProcessBatch(x[2], y[2], S, n)
    ref = O.PortModel.from_file(mf)
    r0 = ref.process(np.concatenate([x[0][0], x[1][0], x[2][0]]))       # slot 0 saw all three calls
    assert float(np.abs(np.concatenate([y[0][0], y[1][0], y[2][0]]) - r0).max()) <= tol
    r9 = O.PortModel.from_file(mf).process(np.concatenate([x[0][9], x[2][9]]))   # slot 9 skipped the middle call
    assert float(np.abs(np.concatenate([y[0][9], y[2][9]]) - r9).max()) <= tol
    # one long call
    xl = (np.random.default_rng(78).uniform(-1, 1, 20000) * amp).astype(np.float32)
    a = _load(na, mf).Process(xl.copy())
    b = _blocks(_load(na, mf), xl, 128)
    assert np.array_equal(a, b)
    assert float(np.abs(a - O.PortModel.from_file(mf).process(xl)).max()) <= tol


def test_errors_are_loud(na, tmp_path):
    g = load_golden(golden_files("syn_a1_nano")[0])
    mf = model_file_for(g, tmp_path)
    m = _load(na, mf, streams=4)
    x = np.zeros((8, 16), dtype=np.float32)
    with pytest.raises(na.NeuralAudioError, match="stream slots"):
        m.ProcessBatch(x, np.empty_like(x), 8, 16)
    with pytest.raises(na.NeuralAudioError):
        na.NeuralModelLoader().CreateFromFile(str(tmp_path / "missing.nam"))
