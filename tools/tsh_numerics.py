"""CPU emulation of the fp16-pair ("2xFP16 split", three products) arithmetic considered for the tcgen05 WaveNet
kernel: every A operand is x ~ h1 + h2 with h1 = rn_f16(x), h2 = rn_f16(x - h1); every weight W ~ W1 + W2 likewise;
D += h1 W1 + h2 W1 + h1 W2 with fp32 accumulation (kind::f16 MMA, K = 16).  The residual stream stays in an fp32
accumulator, the history rings hold the (h1, h2) pairs.  Compared against the committed golden vectors of the reference.
Usage: python tools/tsh_numerics.py [model.nam golden.npz]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ts_numerics as T  # noqa: E402

f32 = np.float32
f16 = np.float16
STATS = {"amax": 0.0, "ovf": 0}


def split_h(x):
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        h1 = x.astype(f16)
        h2 = (x - h1.astype(f32)).astype(f16)
    STATS["amax"] = max(STATS["amax"], float(np.abs(x).max()) if x.size else 0.0)
    STATS["ovf"] += int(np.isinf(h1).sum())
    return h1.astype(np.float64), h2.astype(np.float64)


def mma_h(acc, a, w, variant=None):
    a1, a2 = split_h(a)
    w1, w2 = split_h(w)
    K = a.shape[1]
    for k0 in range(0, K, 16):
        s = slice(k0, min(K, k0 + 16))
        for (p, q) in ((a1, w1), (a2, w1), (a1, w2)):
            acc = (acc.astype(np.float64) + p[:, s] @ q[s, :]).astype(f32)
    return acc


def main():
    root = T.ROOT
    cases = [("BossWN-standard.nam", "ref_BossWN_standard.npz"), ("namcore_wavenet_a1_standard.nam", "ref_namcore_wavenet_a1_standard.npz")]
    if len(sys.argv) > 2:
        cases = [(sys.argv[1], sys.argv[2])]
    n = int(os.environ.get("N", "4096"))
    for m, gname in cases:
        mpath = m if os.path.isabs(m) else os.path.join(root, "oracle", "_ref", "models", m)
        gpath = gname if os.path.isabs(gname) else os.path.join(root, "tests", "golden", gname)
        model = json.load(open(mpath))
        g = np.load(gpath)
        x, y = g["x"][:n], g["y"][:n]
        for amp in (1.0,):
            T.mma3 = T.mma3  # 3xTF32 as the round-1 kernel
            o3 = T.run(model, x * amp, "fma", True)
            save = T.mma3
            T.mma3 = mma_h
            STATS["amax"] = 0.0; STATS["ovf"] = 0
            oh = T.run(model, x * amp, "fma", True)
            T.mma3 = save
            oe = T.run(model, x * amp, "none", False, exact=True)
            if amp == 1.0:
                print("%s: golden vs fp64-contraction %.3g | 3xTF32 %.3g | fp16 pairs %.3g (max |operand| %.3g, overflows %d)" % (
                    m, np.abs(oe - y).max(), np.abs(o3 - y).max(), np.abs(oh - y).max(), STATS["amax"], STATS["ovf"]))


if __name__ == "__main__":
    main()
