"""Single-stream latency of Process() (BASELINE.json cfg 1: one stream, 128-frame buffers): where a call's time goes.
  host     : Process(host in, host out)                      -- what ModelTest's BenchModel measures
  dev+sync : ProcessBatch(device ptrs, S=1) + Synchronize()  -- launch + kernel + sync, no copies
  dev async: 200 back-to-back device calls / 200             -- kernel-bound time per call
Usage: python tools/latency_probe.py [model files ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import neuralaudio_b200 as na


def main():
    files = sys.argv[1:] or [os.path.join(ROOT, "oracle", "_ref", "models", f) for f in
                             ("BossWN-nano.nam", "BossWN-standard.nam", "BossLSTM-1x16.nam", "BossWN-a2.nam")]
    n = 128
    for f in files:
        m = na.NeuralModelLoader().CreateFromFile(f)
        x = np.zeros(n, dtype=np.float32)
        y = np.zeros(n, dtype=np.float32)
        for _ in range(50):
            m.Process(x, y)
        t0 = time.perf_counter()
        for _ in range(2048):
            m.Process(x, y)
        host = (time.perf_counter() - t0) / 2048
        xd = torch.zeros(1, n, device="cuda")
        yd = torch.zeros(1, n, device="cuda")
        for _ in range(50):
            m.ProcessBatch(xd, yd, 1, n)
        m.Synchronize()
        t0 = time.perf_counter()
        for _ in range(1000):
            m.ProcessBatch(xd, yd, 1, n)
            m.Synchronize()
        devsync = (time.perf_counter() - t0) / 1000
        t0 = time.perf_counter()
        for _ in range(1000):
            m.ProcessBatch(xd, yd, 1, n)
        m.Synchronize()
        devasync = (time.perf_counter() - t0) / 1000
        print("%-24s host %.1f us/call (%.1f xRT)   dev+sync %.1f us   dev async %.1f us" %
              (os.path.basename(f), host * 1e6, n / 48000.0 / host, devsync * 1e6, devasync * 1e6))


if __name__ == "__main__":
    main()
