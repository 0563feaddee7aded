"""Randomised parity sweep on the GPU: random call sizes (1..300 frames), stream counts, layouts and host/device buffers over
every architecture family, each checked stream against its own oracle instance fed the concatenated input.
Usage: python tools/fuzz_parity.py [seconds] [seed]  (prints one line per trial and a summary)"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as C
import neuralaudio_b200 as na
from oracle import oracle as O

NAMES = ["syn_a1_standard.", "syn_a1_lite", "syn_a1_feather", "syn_a1_nano.", "syn_a2_full", "syn_a2_lite", "syn_dyn_20x10", "syn_dyn_7x3",
         "syn_dyn_single6_k2", "syn_dyn_16x16_k5", "syn_lstm_1x16", "syn_lstm_2x8", "syn_lstm_1x24", "syn_lstm_2x12", "syn_dyn_lstm_3x18",
         "syn_dyn_lstm_1x40", "syn_a1_standard_sr96000", "syn_a1_nano_sr96000",
         "syn_a2_full_sr96000", "syn_a2_lite_sr96000", "syn_dyn_3arrays", "syn_dyn_4arrays", "syn_dyn_48x24", "syn_dyn_single40_k3",
         "syn_lstm_1x16+tc", "syn_lstm_2x8+tc", "syn_lstm_1x24+tc", "syn_lstm_2x12+tc", "syn_lstm_2x16+tc", "syn_dyn_lstm_2x32+tc", "ref_BossLSTM_1x16+tc"]
# (+tc: the tcgen05 LSTM kernel forced - its automatic choice starts at thousands of stream slots; stream counts up to 300 for it)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    worst = {"wavenet": 0.0, "lstm": 0.0}
    trials = 0
    fails = 0
    with tempfile.TemporaryDirectory() as tmp:
        while time.time() < t_end:
            name, _, force_tc = NAMES[trials % len(NAMES)].partition("+")
            g = C.load_golden(C.golden_files(name)[0])
            mf = C.model_file_for(g, tmp)
            sr = C.external_sample_rate_of(g)
            lstm = C.is_lstm_case(g)
            S = int(rng.integers(1, 601 if force_tc else 41))
            calls = int(rng.integers(2, 7))
            sizes = [int(rng.integers(1, 301)) for _ in range(calls)]
            if rng.random() < 0.3:
                sizes[int(rng.integers(0, calls))] = 128
            layout = int(rng.integers(0, 2))
            on_device = bool(rng.integers(0, 2))
            prev_kernel = na.set_option("lstm_kernel", 4 if force_tc else 0)
            prev_sets = na.set_option("lstm_tc_sets", int(rng.integers(1, 3)) if force_tc else 0)
            try:
                ld = na.NeuralModelLoader()
                ld.SetExternalSampleRate(sr)
                ld.SetDefaultNumStreams(S)
                m = ld.CreateFromFile(mf)
            finally:
                na.set_option("lstm_kernel", prev_kernel)
                na.set_option("lstm_tc_sets", prev_sets)
            amp = 0.5 if lstm else 1.0
            xs = [(rng.uniform(-1, 1, (S, n)) * amp).astype(np.float32) for n in sizes]
            ys = []
            for x in xs:
                n = x.shape[1]
                xin = np.ascontiguousarray(x.T) if layout == 1 else x
                if on_device:
                    xd = torch.from_numpy(xin).cuda()
                    yd = torch.empty_like(xd)
                    m.ProcessBatch(xd, yd, S, n, layout)
                    m.Synchronize()
                    y = yd.cpu().numpy()
                else:
                    y = np.empty_like(xin)
                    m.ProcessBatch(xin, y, S, n, layout)
                ys.append(y.T if layout == 1 else y)
            err = 0.0
            for s in sorted({0, S // 2, S - 1}):
                ref = O.PortModel.from_file(mf, external_sample_rate=sr).process(np.concatenate([x[s] for x in xs]))
                got = np.concatenate([y[s] for y in ys])
                err = max(err, float(np.abs(ref - got).max()))
            tol = C.LSTM_TOL if lstm else C.WAVENET_TOL
            kind = "lstm" if lstm else "wavenet"
            worst[kind] = max(worst[kind], err)
            ok = err <= tol
            fails += 0 if ok else 1
            print("%-26s S=%-3d sizes=%-28s layout=%d %s max-abs %.3g %s" % (g["name"] + ("+tc" if force_tc else ""), S, sizes, layout, "device" if on_device else "host  ", err, "ok" if ok else "FAIL"))
            trials += 1
    print("trials %d, failures %d, worst WaveNet %.3g (tol %.0e), worst LSTM %.3g (tol %.0e), seed %d" %
          (trials, fails, worst["wavenet"], C.WAVENET_TOL, worst["lstm"], C.LSTM_TOL, seed))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
