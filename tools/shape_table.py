"""Device-resident batch throughput of every architecture the loader dispatches, one line per shape (synthetic weights of the
committed golden vectors; 4096 streams x 128 frames per call, A2: 256 frames, LSTM: 8192 streams).
Usage: python tools/shape_table.py > profiles/<round>_shape_table.txt"""
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

SHAPES = ["syn_a1_standard.", "syn_a1_lite", "syn_a1_feather", "syn_a1_nano.", "syn_a2_full", "syn_a2_lite", "syn_a1_standard_sr96000",
          "syn_dyn_20x10", "syn_dyn_16x16_k5", "syn_dyn_7x3", "syn_dyn_single6_k2",
          "syn_dyn_3arrays", "syn_dyn_48x24",
          "syn_lstm_1x16", "syn_lstm_1x24", "syn_lstm_2x8", "syn_lstm_2x12", "syn_lstm_2x16", "syn_dyn_lstm_2x32", "syn_dyn_lstm_3x18", "syn_dyn_lstm_1x40",
          "syn_lstm_1x16@32768", "syn_lstm_1x24@32768", "syn_lstm_2x16@32768"]


def main():
    with tempfile.TemporaryDirectory() as tmp:
        for name in SHAPES:
            name, _, big = name.partition("@")
            g = C.load_golden(C.golden_files(name)[0])
            mf = C.model_file_for(g, tmp)
            sr = C.external_sample_rate_of(g)
            lstm = C.is_lstm_case(g)
            S = int(big) if big else 8192 if lstm else 4096
            n = 256 if "a2_" in name else 128
            d = na.describe_model_file(mf, sr)
            kernel = (d.get("kernel_32768_streams") if big else d.get("kernel")) or ("lstm" if lstm else "")
            if "dyn_20x10" in name or "16x16" in name or "single6" in name or "dyn_3arrays" in name or "dyn_48x24" in name or ("dyn_lstm" in name and "2x32" not in name):
                S //= 8      # run-time-shaped kernels: correctness paths, a smaller batch keeps the table quick
            ld = na.NeuralModelLoader()
            ld.SetExternalSampleRate(sr)
            ld.SetDefaultNumStreams(S)
            m = ld.CreateFromFile(mf)
            x = (torch.rand((4, S, n), device="cuda") * 2 - 1) * (0.5 if lstm else 1.0)
            y = torch.empty_like(x)
            stream = torch.cuda.ExternalStream(m.GetCudaStream())
            for k in range(4):
                m.ProcessBatch(x[k], y[k], S, n)
            m.Synchronize()
            steps = 20
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(steps):
                m.ProcessBatch(x[k % 4], y[k % 4], S, n)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            print("%-26s %-28s %5d streams x %3d frames: %8.1f us/call  %7.3f Gsamples/s  state %6.1f KB/stream" %
                  (g["name"], kernel, S, n, ms * 1e3, S * n / ms / 1e6, m.GetStateBytesPerStream() / 1024.0))


if __name__ == "__main__":
    main()
