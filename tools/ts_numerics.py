"""CPU emulation of the arithmetic the tcgen05 'TS' WaveNet kernel performs (3xTF32 split products with hardware
truncation, bias / mix-in / residual / head folded into MMAs, Horner-form FastMath tanh), to choose the split
variant before spending GPU time.  Compares against a committed golden vector of the reference.
Usage: python tools/ts_numerics.py [model.nam golden.npz]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
f32 = np.float32


def trunc_tf32(x):
    return (np.asarray(x, dtype=f32).view(np.uint32) & np.uint32(0xFFFFE000)).view(f32)


def rn_tf32(x):
    u = np.asarray(x, dtype=f32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(f32)


def split_a(x, variant):
    x = np.asarray(x, dtype=f32)
    hi = trunc_tf32(x)
    if variant == "tweak":      # exact lo, + half an 11-bit ulp on the bit pattern (what the round-1 kernel does)
        lo = (x - hi).astype(f32)
        lo = (lo.view(np.uint32) + np.uint32(0x1000)).view(f32)
    elif variant == "none":
        lo = (x - hi).astype(f32)
    elif variant == "fma":      # lo = fma(x, 1 + 2^-23, -hi): one instruction, de-biases on average
        lo = (x.astype(np.float64) * (1.0 + 2.0 ** -23) - hi.astype(np.float64)).astype(f32)
    elif variant == "fma2":
        lo = (x.astype(np.float64) * (1.0 + 2.0 ** -24) - hi.astype(np.float64)).astype(f32)
    else:
        raise ValueError(variant)
    return x, lo   # hi is fed raw (hardware truncates)


def split_b(w):
    w = np.asarray(w, dtype=f32)
    hi = rn_tf32(w)
    return hi, (w - hi).astype(f32)


def mma3(acc, a, w, variant):
    """acc[T,N] (f32) += a[T,K] @ w[K,N] as 3xTF32 in K-steps of 8 with fp32 accumulator rounding per MMA."""
    ah, al = split_a(a, variant)
    wh, wl = split_b(w)
    ah = trunc_tf32(ah).astype(np.float64); al = trunc_tf32(al).astype(np.float64)
    wh = trunc_tf32(wh).astype(np.float64); wl = trunc_tf32(wl).astype(np.float64)
    K = a.shape[1]
    for k0 in range(0, K, 8):
        s = slice(k0, min(K, k0 + 8))
        for (p, q) in ((ah, wh), (al, wh), (ah, wl)):
            acc = (acc.astype(np.float64) + p[:, s] @ q[s, :]).astype(f32)
    return acc


def fast_tanh_ref(x):
    x = x.astype(f32)
    ax = np.abs(x); x2 = x * x
    num = x * (f32(2.45550750702956) + f32(2.45550750702956) * ax + (f32(0.893229853513558) + f32(0.821226666969744) * ax) * x2)
    den = f32(2.44506634652299) + (f32(2.44506634652299) + x2) * np.abs(x + f32(0.814642734961073) * x * ax)
    return (num / den).astype(f32)


def fast_tanh_horner(x):
    x = x.astype(f32)
    a = np.abs(x)
    c0 = f32(2.45550750702956); c1 = f32(0.893229853513558); c2 = f32(0.821226666969744)
    c3 = f32(2.44506634652299); c4 = f32(0.814642734961073)
    n = (c2 * a + c1).astype(f32); n = (n * a + c0).astype(f32); n = (n * a + c0).astype(f32)
    d = (c4 * a + f32(1.0)).astype(f32); d = (d * a + f32(c3 * c4)).astype(f32); d = (d * a + c3).astype(f32); d = (d * a + c3).astype(f32)
    return ((x * n).astype(f32) * (f32(1.0) / d).astype(f32)).astype(f32)


def run(model, x, variant, horner, exact=False):
    cfg = model["config"]
    w = np.asarray(model["weights"], dtype=f32)
    pos = 0

    def take(n):
        nonlocal pos
        v = w[pos:pos + n]; pos += n
        return v

    arrays = []
    for A in cfg["layers"]:
        C, inC, H, K = A["channels"], A["input_size"], A["head_size"], A["kernel_size"]
        re = take(C * inC).reshape(C, inC)
        layers = []
        for d in A["dilations"]:
            conv = take(C * C * K).reshape(C, C, K)   # [out][in][k]
            cb = take(C); mix = take(C)
            one = take(C * C).reshape(C, C); ob = take(C)
            layers.append((d, conv, cb, mix, one, ob))
        hw = take(H * C).reshape(H, C)
        hb = take(H) if A["head_bias"] else np.zeros(H, f32)
        arrays.append((C, inC, H, K, re, layers, hw, hb))
    head_scale = take(1)[0]
    RF = 4092
    pad = RF + 128
    xin = np.concatenate([np.zeros(pad, f32), x.astype(f32)])
    T = xin.size
    cond = xin[:, None]

    def mm(acc, a, wt):
        if exact:
            return (acc.astype(np.float64) + a.astype(np.float64) @ wt.astype(np.float64)).astype(f32)
        return mma3(acc, a, wt, variant)

    ones = np.ones((T, 1), f32)
    layer_in = cond
    head_in = None
    for ai, (C, inC, H, K, re, layers, hw, hb) in enumerate(arrays):
        xcur = mm(np.zeros((T, C), f32), layer_in, re.T.copy())
        head_acc = np.zeros((T, H), f32) if head_in is None else mm(np.zeros((T, H), f32), head_in, hw.T.copy())
        last_array = ai + 1 == len(arrays)
        for li, (d, conv, cb, mix, one, ob) in enumerate(layers):
            z = np.zeros((T, C), f32)
            for k in range(K):
                D = (K - 1 - k) * d
                xs = np.concatenate([np.repeat(xcur[:1], D, axis=0), xcur[:T - D]]) if D > 0 else xcur
                z = mm(z, xs, conv[:, :, k].T.copy())
            # bias + mix-in folded as one more MMA: A = [cond, 1], B = [mix; bias]
            z = mm(z, np.concatenate([cond, ones], axis=1), np.stack([mix, cb]))
            zt = fast_tanh_horner(z) if horner else fast_tanh_ref(z)
            head_acc = mm(head_acc, zt, hw.T.copy())
            if not (last_array and li == len(layers) - 1):
                xcur = mm(xcur, zt, one.T.copy())
                xcur = mm(xcur, ones, ob[None, :])
        if last_array:
            out = (head_acc[:, 0] + hb[0]).astype(f32) * f32(head_scale)
        else:
            layer_in = xcur
            head_in_next = head_acc   # pre-bias head outputs feed the next array's head sum (no bias in A1 array 0)
            head_in = (head_in_next + hb[None, :]).astype(f32)
    return out[pad:]


def main():
    mpath = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "models", "BossWN-standard.nam")
    gpath = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "ref_BossWN_standard.npz")
    model = json.load(open(mpath))
    g = np.load(gpath)
    x, y = g["x"], g["y"]
    n = int(os.environ.get("N", "4096"))
    x, y = x[:n], y[:n]
    o = run(model, x, "none", False, exact=True)
    print("fp64-contraction restatement vs golden: max-abs %.3g" % np.abs(o - y).max())
    for variant in ("tweak", "none", "fma", "fma2"):
        for horner in (False, True):
            o = run(model, x, variant, horner)
            e = o - y
            print("3xTF32 split=%-6s horner=%d : max-abs %.3g  mean err %.3g  rms %.3g" % (variant, horner, np.abs(e).max(), e.mean(), np.sqrt((e * e).mean())))


if __name__ == "__main__":
    main()
