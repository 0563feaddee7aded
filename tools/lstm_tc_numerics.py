"""CPU emulation of the arithmetic a tensor-core LSTM kernel would perform: the gate
pre-activations W . [x; h] as 3xTF32 split products (A = [x | h] truncated by the hardware, B = weights rounded to tf32,
fp32 accumulation per K-step of 8), everything else (FastMath activations with the IEEE quotient, c / h updates, head dot)
in fp32 as today.  Compares against the committed golden vectors of the reference to decide, before any GPU work, whether
the formulation stays inside the LSTM tolerance (5e-5): an LSTM feeds its own output back, so split errors could compound.
Usage: python tools/lstm_tc_numerics.py [golden names ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest as C

f32 = np.float32


def trunc_tf32(x):
    return (np.asarray(x, dtype=f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def rn_tf32(x):
    u = np.asarray(x, dtype=f32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)


def gates_3xtf32(W, s):
    """W [4H][K] fp32, s [K] fp32 -> W . s as the tensor core would accumulate it (K-steps of 8, fp32 accumulator)."""
    wh = rn_tf32(W)
    wl = (W - wh).astype(f32)
    sh = trunc_tf32(s)
    sl = (s - sh).astype(f32)   # exact low part (the WaveNet kernel de-biases it; a second-order effect)
    acc = np.zeros(W.shape[0], f32)
    K = W.shape[1]
    for k0 in range(0, K, 8):
        sl_ = slice(k0, min(K, k0 + 8))
        for a, b in ((sh, wh), (sl, wh), (sh, wl)):
            acc = (acc.astype(np.float64) + trunc_tf32(b[:, sl_]).astype(np.float64) @ trunc_tf32(a[sl_]).astype(np.float64)).astype(f32)
    return acc


def gates_fp32(W, s):
    acc = np.zeros(W.shape[0], f32)
    for j in range(W.shape[1]):
        acc = (acc + W[:, j] * s[j]).astype(f32)      # (separate rounding of the product: a plain fp32 restatement)
    return acc


def fast_tanh(x):
    x = x.astype(f32)
    ax = np.abs(x)
    x2 = x * x
    num = x * (f32(2.45550750702956) + f32(2.45550750702956) * ax + (f32(0.893229853513558) + f32(0.821226666969744) * ax) * x2)
    den = f32(2.44506634652299) + (f32(2.44506634652299) + x2) * np.abs(x + f32(0.814642734961073) * x * ax)
    return (num / den).astype(f32)


def fast_sigmoid(x):
    return (f32(0.5) * (fast_tanh((x * f32(0.5)).astype(f32)) + f32(1.0))).astype(f32)


def run(model, weights, x, gate_fn, prewarm=2048):
    cfg = model["config"]
    L, H = int(cfg["num_layers"]), int(cfg["hidden_size"])
    w = np.asarray(weights, dtype=f32)
    p = 0
    layers = []
    for l in range(L):
        I = 1 if l == 0 else H
        W = w[p:p + 4 * H * (I + H)].reshape(4 * H, I + H); p += 4 * H * (I + H)
        b = w[p:p + 4 * H]; p += 4 * H
        h0 = w[p:p + H].copy(); p += H
        c0 = w[p:p + H].copy(); p += H
        layers.append([W, b, h0, c0])
    hw = w[p:p + H]; p += H
    hb = w[p]
    xin = np.concatenate([np.zeros(prewarm, f32), x.astype(f32)])
    out = np.empty(xin.size, f32)
    for t in range(xin.size):
        inp = np.array([xin[t]], f32)
        for Ly in layers:
            W, b, h, c = Ly
            s = np.concatenate([inp, h]).astype(f32)
            g = (gate_fn(W, s) + b).astype(f32)
            i, f, gg, o = g[0:H], g[H:2 * H], g[2 * H:3 * H], g[3 * H:4 * H]
            c = (fast_sigmoid(f) * c + fast_sigmoid(i) * fast_tanh(gg)).astype(f32)
            h = (fast_sigmoid(o) * fast_tanh(c)).astype(f32)
            Ly[2], Ly[3] = h, c
            inp = h
        out[t] = f32(np.dot(hw.astype(np.float64), inp.astype(np.float64)) + hb)
    return out[prewarm:]


def main():
    names = sys.argv[1:] or ["syn_lstm_1x16", "syn_lstm_1x24", "syn_lstm_2x12", "syn_lstm_2x16"]
    n = int(os.environ.get("N", "3000"))
    for name in names:
        g = C.load_golden(C.golden_files(name)[0])
        x, y = g["x"][:n], g["y"][:n]
        e32 = float(np.abs(run(g["model"], g["weights"], x, gates_fp32) - y).max())
        etc = float(np.abs(run(g["model"], g["weights"], x, gates_3xtf32) - y).max())
        print("%-16s %d samples: plain fp32 restatement max-abs %.3g | 3xTF32 gates max-abs %.3g | tolerance %.0e" % (name, n, e32, etc, C.LSTM_TOL))


if __name__ == "__main__":
    main()
