"""GPU: the BASELINE.json configurations at their FULL sizes, the multi-GPU load paths and batch-level quality switching.

Full sizes cannot be replayed stream by stream through the oracle, so they are checked through size-independent
properties: stream s carries input pattern s % 8, hence all streams of a pattern must be
bit-identical wherever the kernel's grid / wave placement puts them (first CTA round, last partial wave, ...), and one
stream per pattern is compared sample by sample with the oracle."""
import numpy as np
import pytest

from conftest import WAVENET_TOL, LSTM_TOL, golden_files, load_golden, model_file_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def _load(na, mf, quality=1.0, streams=1, prewarm=True, device=None):
    ld = na.NeuralModelLoader()
    ld.SetDefaultQualityScaleFactor(quality)
    ld.SetDefaultNumStreams(streams)
    if device is not None:
        ld.SetDevice(device)
    return ld.CreateFromFile(mf, prewarm)


def _patterned_full_size(na, O, tmp_path, name, S, n, calls, tol, amplitude=1.0):
    import torch
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    rng = np.random.default_rng(23)
    base = (rng.uniform(-1, 1, (calls, 8, n)) * amplitude).astype(np.float32)
    x = np.tile(base, (1, S // 8, 1))           # stream s carries pattern s % 8
    m = _load(na, mf, streams=S)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    for k in range(calls):
        m.ProcessBatch(xd[k], yd[k], S, n)
    m.Synchronize()
    y = yd.cpu().numpy()
    worst = 0.0
    for p in range(8):
        # every stream of the pattern, in whichever CTA round or partial wave it ran: bit-identical
        assert np.array_equal(y[:, p::8, :], np.broadcast_to(y[:, p:p + 1, :], y[:, p::8, :].shape)), "pattern %d differs between streams" % p
        ys = O.PortModel.from_file(mf).process(np.ascontiguousarray(base[:, p, :]).reshape(-1))
        worst = max(worst, float(np.abs(ys - y[:, S - 8 + p, :].reshape(-1)).max()))   # probe the LAST streams of the batch
    assert worst <= tol, "max-abs vs oracle %.3g" % worst
    return m, mf


def test_full_size_cfg3_lstm_8192x128(na, O, tmp_path):
    """BASELINE.json cfg 3: NAM LSTM 1x16, 8192 streams x 128 frames (1.7 waves of the gate-rows-in-registers kernel)."""
    _patterned_full_size(na, O, tmp_path, "syn_lstm_1x16", 8192, 128, 6, LSTM_TOL, amplitude=0.5)


def test_full_size_lstm_tensor_core_kernel(na, O, tmp_path):
    """The batches the automatic choice gives to the tcgen05 LSTM kernel: 2x16 at 8192 streams (64-stream CTAs, one per SM) and
    1x16 at 16384 streams (two CTAs per SM) and at 32768 (two 128-stream sets per CTA)."""
    assert na.describe_model_file(model_file_for(load_golden(golden_files("syn_lstm_2x16")[0]), tmp_path))["kernel"] == "lstm_tcgen05_gates"
    _patterned_full_size(na, O, tmp_path, "syn_lstm_2x16", 8192, 128, 4, LSTM_TOL, amplitude=0.5)
    _patterned_full_size(na, O, tmp_path, "syn_lstm_1x16", 16384, 128, 4, LSTM_TOL, amplitude=0.5)
    _patterned_full_size(na, O, tmp_path, "syn_lstm_1x16", 32768, 128, 3, LSTM_TOL, amplitude=0.5)    # two 128-stream sets per CTA


def test_full_size_cfg5_a2_full_4096x256(na, O, tmp_path):
    """BASELINE.json cfg 5: NAM A2 'Full' (8 channels), 4096 streams x 256 frames (two 128-frame passes per call)."""
    m, _ = _patterned_full_size(na, O, tmp_path, "syn_a2_full", 4096, 256, 4, WAVENET_TOL)
    assert m.GetStateBytesPerStream() >= 203072


def test_full_size_cfg2_a1_standard_last_wave(na, O, tmp_path):
    """cfg 2 again with the probes on the batch's last streams (the persistent kernel's final, partly filled round)."""
    _patterned_full_size(na, O, tmp_path, "syn_a1_standard.", 4096, 128, 5, WAVENET_TOL)


def test_batch_quality_switch_keeps_every_slots_state(na, O, tmp_path):
    """SetQualityScaleFactor with S > 1 stream slots: each resident sub-model of the A2 container keeps its own per-slot state
    (CompositeModel.h:94-118: the inactive sub-model simply does not advance), checked per stream against the oracle."""
    p = O.model_path("BossWN-a2.nam")
    if p is None:
        pytest.skip("fixture not staged")
    S, n = 9, 128
    rng = np.random.default_rng(31)
    plan = [(1.0, 3), (0.3, 4), (1.0, 2), (0.0, 2)]      # (quality, calls)
    m = _load(na, p, 1.0, streams=S)
    refs = [O.PortModel.from_file(p, quality=1.0) for _ in range(S)]
    worst = 0.0
    for q, calls in plan:
        m.SetQualityScaleFactor(q)
        for r in refs:
            r.set_quality(q)
        for _ in range(calls):
            x = rng.uniform(-1, 1, (S, n)).astype(np.float32)
            y = np.empty_like(x)
            m.ProcessBatch(x, y, S, n)
            for s in range(S):
                worst = max(worst, float(np.abs(refs[s].process(x[s]) - y[s]).max()))
    assert worst <= WAVENET_TOL, worst


@pytest.mark.parametrize("name", ["syn_a1_standard.", "syn_lstm_1x16", "syn_a2_full"])
def test_device_blob_carries_weights_and_prewarmed_state(na, name, tmp_path):
    """NA_GetDeviceBlob + NA_ResetStreams (the caller-driven multi-GPU load): copying a prewarmed model's
    [packed weights | state template] blob into a model built WITHOUT prewarm makes the two bit-identical."""
    from cuda.bindings import runtime as cudart
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    S, n = 5, 128
    a = _load(na, mf, streams=S, prewarm=True)
    b = _load(na, mf, streams=S, prewarm=False)
    pa, na_bytes = a.GetDeviceBlob()
    pb, nb_bytes = b.GetDeviceBlob()
    assert na_bytes == nb_bytes and na_bytes > 0
    a.Synchronize(); b.Synchronize()
    err, = cudart.cudaMemcpy(pb, pa, na_bytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    b.ResetStreams()
    x = np.random.default_rng(3).uniform(-1, 1, (3, S, n)).astype(np.float32)
    ya, yb = np.empty_like(x), np.empty_like(x)
    for k in range(3):
        a.ProcessBatch(x[k], ya[k], S, n)
        b.ProcessBatch(x[k], yb[k], S, n)
    assert np.array_equal(ya, yb)


def test_broadcast_model_single_rank(na, tmp_path):
    """NA_BroadcastModel on a one-rank communicator (the library's own ncclCommInitRank + ncclBroadcast): bytes = the blob,
    outputs unchanged.  The multi-rank case runs in bench.py under torchrun and in the sharded test below."""
    g = load_golden(golden_files("syn_a1_standard.")[0])
    mf = model_file_for(g, tmp_path)
    try:
        uid = na.nccl_get_unique_id()
    except na.NeuralAudioError as e:
        pytest.skip("NCCL not loadable here: %s" % e)
    comm = na.NcclComm(1, 0, uid, 0)
    assert comm.nranks == 1
    S, n = 4, 128
    a = _load(na, mf, streams=S)
    b = _load(na, mf, streams=S)
    nbytes = b.BroadcastModel(comm, 0)
    assert nbytes == b.GetDeviceBlob()[1]
    x = np.random.default_rng(5).uniform(-1, 1, (S, n)).astype(np.float32)
    ya, yb = np.empty_like(x), np.empty_like(x)
    a.ProcessBatch(x, ya, S, n)
    b.ProcessBatch(x, yb, S, n)
    assert np.array_equal(ya, yb)
    comm.close()


@pytest.mark.parametrize("name,S", [("syn_a1_standard.", 300), ("syn_lstm_1x16", 257), ("syn_a2_full", 64)])
def test_sharded_model_equals_single_device(na, name, S, tmp_path):
    """shard outputs equal the single-GPU outputs bit for bit.  One host process, every visible
    GPU (at least the one): NA_CreateModelSharded builds the model on each device, broadcasts device 0's
    [weights | prewarmed state] with ONE grouped ncclBroadcast and fans NA_ProcessBatch out in contiguous stream blocks."""
    import torch
    ndev = min(na.device_count(), 8)
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    n, calls = 128, 4
    ld = na.NeuralModelLoader()
    ld.SetDefaultNumStreams(S)
    sharded = ld.CreateShardedFromFile(mf, list(range(ndev)))
    assert sharded.GetNumShards() == ndev and sharded.GetNumStreams() == S
    if ndev > 1:
        assert sharded.GetBroadcastBytes() > 0
    single = _load(na, mf, streams=S, device=0)
    x = torch.from_numpy(np.random.default_rng(9).uniform(-0.7, 0.7, (calls, S, n)).astype(np.float32)).pin_memory()
    ys = torch.empty_like(x).pin_memory()
    y1 = torch.empty_like(x).pin_memory()
    for k in range(calls):
        sharded.ProcessBatch(x[k], ys[k], S, n)
        single.ProcessBatch(x[k], y1[k], S, n)
    assert torch.equal(ys, y1)
    # a partial batch (fewer streams than slots) still lands every stream on its own device's slot
    part = S - S // 3
    sharded.ProcessBatch(x[0][:part].contiguous().pin_memory(), ys[0][:part], part, n)
    single.ProcessBatch(x[0][:part].contiguous().pin_memory(), y1[0][:part], part, n)
    assert torch.equal(ys[0][:part], y1[0][:part])


@pytest.mark.parametrize("name,S,n", [("syn_a1_standard.", 4096, 128), ("syn_lstm_1x16", 8192, 128), ("syn_a1_nano.", 1500, 64)])
def test_blocking_host_call_is_sliced_and_exact(na, name, S, n, tmp_path):
    """The drop-in blocking NA_ProcessBatch with HOST pointers pipelines slices of the batch (copy-in | kernels | copy-out):
    pinned and pageable buffers must give bit-identical results to the device-pointer path, call after call."""
    import torch
    g = load_golden(golden_files(name)[0])
    mf = model_file_for(g, tmp_path)
    calls = 3
    x = torch.from_numpy(np.random.default_rng(41).uniform(-0.6, 0.6, (calls, S, n)).astype(np.float32))
    ref = _load(na, mf, streams=S)
    xd = x.cuda()
    yd = torch.empty_like(xd)
    for k in range(calls):
        ref.ProcessBatch(xd[k], yd[k], S, n)
    ref.Synchronize()
    want = yd.cpu()
    pinned = _load(na, mf, streams=S)
    xp, yp = x.pin_memory(), torch.empty_like(x).pin_memory()
    pageable = _load(na, mf, streams=S)
    xq, yq = x.numpy().copy(), np.empty((calls, S, n), dtype=np.float32)
    for k in range(calls):
        pinned.ProcessBatch(xp[k], yp[k], S, n)
        pageable.ProcessBatch(xq[k], yq[k], S, n)
    assert torch.equal(yp, want)
    assert np.array_equal(yq, want.numpy())
