"""Tensor-core LSTM kernel (lstm_kernel = 4) against the other LSTM kernels and the oracle, with device timings.
Per shape: S streams x 128 frames, a few calls with odd sizes; max-abs between kernels, max-abs of probed streams vs the oracle
port, microseconds per 128-frame call.  Usage: python tools/lstm_tc_check.py [S]"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as C
import neuralaudio_b200 as na
from oracle import oracle as O

SHAPES = ["syn_lstm_1x16", "syn_lstm_1x24", "syn_lstm_2x8", "syn_lstm_2x12", "syn_lstm_2x16", "syn_dyn_lstm_2x32", "syn_lstm_1x8",
          "ref_BossLSTM_1x16", "ref_BossLSTM_2x8"]


def run(mf, S, xs, kern, time_it):
    prev = na.set_option("lstm_kernel", kern)
    try:
        ld = na.NeuralModelLoader()
        ld.SetDefaultNumStreams(S)
        m = ld.CreateFromFile(mf)
        ys = []
        for x in xs:
            y = torch.empty_like(x)
            m.ProcessBatch(x, y, S, x.shape[1])
            ys.append(y)
        m.Synchronize()
        us = 0.0
        if time_it:
            x = xs[0]
            y = torch.empty_like(x)
            stream = torch.cuda.ExternalStream(m.GetCudaStream())
            for _ in range(3):
                m.ProcessBatch(x, y, S, x.shape[1])
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                m.ProcessBatch(x, y, S, x.shape[1])
            e1.record(stream)
            e1.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
        return torch.cat(ys, dim=1).cpu().numpy(), us
    finally:
        na.set_option("lstm_kernel", prev)


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    only = sys.argv[2:] or SHAPES
    with tempfile.TemporaryDirectory() as tmp:
        for name in only:
            files = C.golden_files(name)
            if not files:
                continue
            g = C.load_golden(files[0])
            mf = C.model_file_for(g, tmp)
            if mf is None:
                continue
            gen = torch.Generator(device="cuda").manual_seed(7)
            xs = [((torch.rand((S, n), device="cuda", generator=gen) * 2 - 1) * 0.5).contiguous() for n in (128, 1, 37, 128, 70)]
            y4, us4 = run(mf, S, xs, 4, True)
            y0, us0 = run(mf, S, xs, 0 if S < 512 else 2 if ("1x24" in name or "2x12" in name or "2x16" in name or "2x32" in name) else 1, True)
            worst = 0.0
            for s in sorted({0, S // 2 + 1, S - 1}):
                ref = O.PortModel.from_file(mf).process(np.concatenate([x[s].cpu().numpy() for x in xs]))
                worst = max(worst, float(np.abs(ref - y4[s]).max()))
            print("%-22s S=%5d  tc %7.1f us  other %7.1f us  (x%.2f)  tc-vs-other %.2e  tc-vs-oracle %.2e  finite %s" %
                  (g["name"], S, us4, us0, us0 / max(us4, 1e-9), float(np.abs(y4 - y0).max()), worst, bool(np.isfinite(y4).all())), flush=True)


if __name__ == "__main__":
    main()
