"""Generates the committed golden vectors in tests/golden/*.npz from the UNMODIFIED reference compiled by
oracle/build_ref.sh (oracle/_ref/libna_ref.so).  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

Two families:

* ``ref_<fixture>.npz``  -- the reference's own fixture models (Utils/Models, NAM Core example_models): only the seeded
  input and the reference's output are stored (the model files themselves are CC BY-NC-ND and are NOT committed; they
  are staged, git-ignored, under oracle/_ref/models/).
* ``syn_<arch>.npz``     -- synthetic models of the official architectures with seeded random weights (ours, so the
  weights are committed too); tests rebuild the .nam from them, so these run even where the fixtures are absent.

Protocol per vector: default loader (prewarm on load), white noise U[-1,1) seeded per case, Process() in 128-frame
calls, 8192 samples (2x the A1 receptive field, > the A2 one), plus a 512-sample all-zero tail run on a fresh model
("dc") that pins the prewarm steady state.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as O   # noqa: E402

N = 8192
BLOCK = 128

STD = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]
LITE1 = [1, 2, 4, 8, 16, 32, 64]
LITE2 = [128, 256, 512, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512]


def a1_config(c, h, official_std):
    d0, d1 = (STD, STD) if official_std else (LITE1, LITE2)
    return {"layers": [
        {"input_size": 1, "condition_size": 1, "head_size": h, "channels": c, "kernel_size": 3, "dilations": d0,
         "activation": "Tanh", "gated": False, "head_bias": False},
        {"input_size": c, "condition_size": 1, "head_size": 1, "channels": h, "kernel_size": 3, "dilations": d1,
         "activation": "Tanh", "gated": False, "head_bias": True}], "head": None, "head_scale": 0.02}


def a1_num_weights(cfg):
    n = 1
    for L in cfg["layers"]:
        c, k = L["channels"], L["kernel_size"]
        n += c * L["input_size"] + len(L["dilations"]) * (c * c * k + c + c + c * c + c) + L["head_size"] * c + (L["head_size"] if L["head_bias"] else 0)
    return n


def a2_config(c):
    ks = O.A2_KERNEL_SIZES
    return {"layers": [{"input_size": 1, "condition_size": 1, "head": {"out_channels": 1, "kernel_size": 16, "bias": True},
                        "channels": c, "kernel_sizes": ks, "dilations": O.A2_DILATIONS,
                        "activation": [{"type": "LeakyReLU", "negative_slope": 0.01}] * len(ks),
                        "gating_mode": ["none"] * len(ks), "secondary_activation": [None] * len(ks),
                        "layer1x1": {"active": True, "groups": 1}, "head1x1": {"active": False},
                        "conv_pre_film": {"active": False}, "conv_post_film": {"active": False},
                        "input_mixin_pre_film": {"active": False}, "input_mixin_post_film": {"active": False},
                        "activation_pre_film": {"active": False}, "activation_post_film": {"active": False},
                        "layer1x1_post_film": {"active": False}, "head1x1_post_film": {"active": False},
                        "bottleneck": c, "groups_input": 1, "groups_input_mixin": 1, "slimmable": None}],
            "head": None, "head_scale": 1.0, "in_channels": 1}


def a2_config_namcore(c):
    """The same A2 architecture with every key NAM Core's own parser reads (the fixture's layout, BossWN-a2.nam): needed where the
    reference hands the file to NAM Core (an A2 model on a host at 96 kHz: OversampleNAMConfig, NeuralModel.cpp:92-130,365-380)."""
    cfg = a2_config(c)
    lc = cfg["layers"][0]
    lc["head1x1"] = {"active": False, "out_channels": 1, "groups": 1}
    for k in ("conv_pre_film", "conv_post_film", "input_mixin_pre_film", "input_mixin_post_film", "activation_pre_film",
              "activation_post_film", "layer1x1_post_film", "head1x1_post_film"):
        lc[k] = {"active": False, "shift": True, "groups": 1}
    del cfg["in_channels"]
    return cfg


def a2_oversampled_vectors(tmpdir):
    """A2 models on a host at 96 kHz.  The reference doubles the dilations and the head dilation, its Internal path refuses the
    result and NAM Core runs it; run with NA_REF_NAMCORE=1 so that oracle/_ref/libna_ref_namcore.so (the reference built WITH
    its NAM Core back-end) produces the vectors."""
    made = []
    rng = np.random.default_rng(20261020)
    for j, (name, c) in enumerate([("a2_full", 8), ("a2_lite", 3)]):
        cfg = a2_config_namcore(c)
        w = synth_wavenet_weights(rng, a2_num_weights(c), 1.0 / np.sqrt(6 * c), 1.0)
        case = {"version": "0.7.0", "architecture": "WaveNet", "config": cfg, "weights": w, "sample_rate": 48000,
                "metadata": {"loudness": -12.0, "name": "syn-" + name + "-96k"}}
        path = os.path.join(tmpdir, name + "_nc.nam")
        with open(path, "w") as f:
            f.write(nam_text(case))
        x = np.random.default_rng(555 + j).uniform(-1.0, 1.0, N).astype(np.float32)
        y, dc, info = run_ref(path, x, external_sample_rate=96000)
        meta = {k: v for k, v in case.items() if k != "weights"}
        out = os.path.join(HERE, "syn_%s_sr96000.npz" % name)
        np.savez_compressed(out, x=x, y=y, dc=dc, weights=np.asarray(w, dtype=np.float32), model=json.dumps(meta),
                            info=json.dumps(info), external_sample_rate=np.int32(96000))
        made.append(out)
        os.remove(path)
    return made


def a2_num_weights(c):
    n = c + 1
    for k in O.A2_KERNEL_SIZES:
        n += c * c * k + c + c + c * c + c
    return n + c * 16 + 1


def synth_wavenet_weights(rng, n, scale, head_scale):
    w = rng.uniform(-1.0, 1.0, n).astype(np.float32) * np.float32(scale)
    w[-1] = np.float32(head_scale)
    return w


def synthetic_cases():
    rng = np.random.default_rng(20261017)
    cases = {}
    for name, (c, h, std) in {"a1_standard": (16, 8, True), "a1_lite": (12, 6, False), "a1_feather": (8, 4, False),
                              "a1_nano": (4, 2, False)}.items():
        cfg = a1_config(c, h, std)
        w = synth_wavenet_weights(rng, a1_num_weights(cfg), 1.1 / np.sqrt(3 * c), 1.0)
        cases[name] = {"version": "0.5.4", "architecture": "WaveNet", "config": cfg, "weights": w, "sample_rate": 48000,
                       "metadata": {"loudness": -10.5, "name": "syn-" + name, "input_level_dbu": 9.5}}
    for name, c in {"a2_full": 8, "a2_lite": 3}.items():
        w = synth_wavenet_weights(rng, a2_num_weights(c), 1.3 / np.sqrt(6 * c), 0.5 if c == 8 else 0.1)
        cases[name] = {"version": "0.7.0", "architecture": "WaveNet", "config": a2_config(c), "weights": w, "sample_rate": 48000,
                       "metadata": {"loudness": -12.25}}
    for name, (L, H) in {"lstm_1x16": (1, 16), "lstm_2x8": (2, 8), "lstm_1x24": (1, 24), "lstm_2x12": (2, 12)}.items():
        n = H + 1
        for l in range(L):
            i = 1 if l == 0 else H
            n += 4 * H * (i + H) + 4 * H + 2 * H
        w = rng.uniform(-1.0, 1.0, n).astype(np.float32) * np.float32(2.5 / np.sqrt(H))
        cases[name] = {"version": "0.5.2", "architecture": "LSTM", "config": {"input_size": 1, "hidden_size": H, "num_layers": L},
                       "weights": w, "sample_rate": 48000, "metadata": {"loudness": -15.0}}
    return cases


def dyn_config(arrays):
    """Run-time-shaped A1-style stacks (the reference's dynamic path, WaveNetDynamic.h): arrays = [(channels, kernel, dilations), ...]"""
    layers = []
    for i, (c, k, dil) in enumerate(arrays):
        last = i + 1 == len(arrays)
        layers.append({"input_size": 1 if i == 0 else arrays[i - 1][0], "condition_size": 1, "head_size": 1 if last else arrays[i + 1][0],
                       "channels": c, "kernel_size": k, "dilations": dil, "activation": "Tanh", "gated": False, "head_bias": last})
    return {"layers": layers, "head": None, "head_scale": 0.02}


def dynamic_cases():
    rng = np.random.default_rng(20261018)
    shapes = {"dyn_20x10": [(20, 3, [1, 2, 4, 8, 16, 32, 64]), (10, 3, [128, 1, 2, 4, 8])],
              "dyn_single6_k2": [(6, 2, [1, 3, 9, 27])],
              "dyn_16x16_k5": [(16, 5, [1, 2, 4, 8]), (16, 5, [1, 2, 4, 8])],
              "dyn_7x3": [(7, 3, [1, 2, 4, 8, 16]), (3, 3, [1, 2, 4])],
              # beyond two layer arrays / 32 channels (WaveNetDynamic.h takes any count and width); appended so that the draws of
              # the earlier cases stay what they were
              "dyn_3arrays": [(8, 3, [1, 2, 4, 8, 16]), (6, 3, [1, 2, 4, 8]), (4, 2, [1, 3, 9])],
              "dyn_4arrays": [(5, 2, [1, 2, 4]), (10, 3, [1, 2]), (3, 3, [1, 4, 16]), (2, 2, [1, 2])],
              "dyn_48x24": [(48, 3, [1, 2, 4, 8, 16, 32]), (24, 3, [1, 2, 4, 8])],
              "dyn_single40_k3": [(40, 3, [1, 2, 4, 8, 16, 32, 64, 128])]}
    cases = {}
    for name, arrays in shapes.items():
        cfg = dyn_config(arrays)
        c, k = arrays[0][0], arrays[0][1]
        w = synth_wavenet_weights(rng, a1_num_weights(cfg), 1.1 / np.sqrt(k * c), 1.0)
        cases[name] = {"version": "0.5.4", "architecture": "WaveNet", "config": cfg, "weights": w, "sample_rate": 48000,
                       "metadata": {"loudness": -11.0, "name": "syn-" + name}}
    return cases


def dynamic_lstm_cases():
    """LSTM sizes outside the reference's static list (NeuralModel.cpp:25-38): its dynamic path (LSTMDynamic.h)."""
    rng = np.random.default_rng(20261019)
    cases = {}
    # (lstm_1x8 / lstm_2x16 are static shapes of the reference, added with this batch so that every static LSTM size has a vector)
    for name, (L, H) in {"dyn_lstm_3x18": (3, 18), "dyn_lstm_1x40": (1, 40), "dyn_lstm_4x6": (4, 6), "dyn_lstm_2x32": (2, 32),
                         "lstm_1x8": (1, 8), "lstm_2x16": (2, 16)}.items():
        n = H + 1
        for l in range(L):
            i = 1 if l == 0 else H
            n += 4 * H * (i + H) + 4 * H + 2 * H
        w = rng.uniform(-1.0, 1.0, n).astype(np.float32) * np.float32(2.5 / np.sqrt(H))
        cases[name] = {"version": "0.5.2", "architecture": "LSTM", "config": {"input_size": 1, "hidden_size": H, "num_layers": L},
                       "weights": w, "sample_rate": 48000, "metadata": {"loudness": -15.0}}
    return cases


def nam_text(case):
    d = dict(case)
    d["weights"] = [float(x) for x in np.asarray(case["weights"], dtype=np.float32)]
    return json.dumps(d)


def run_ref(path, x, quality=1.0, external_sample_rate=48000):
    m = O.RefModel(path, quality=quality, external_sample_rate=external_sample_rate)
    y = m.process_blocks(x, BLOCK)
    m.close()
    m = O.RefModel(path, quality=quality, external_sample_rate=external_sample_rate)
    dc = m.process_blocks(np.zeros(512, dtype=np.float32), BLOCK)
    info = dict(static=m.is_static(), rf=m.receptive_field(), sample_rate=m.sample_rate(), in_adj=m.input_adjust(),
                out_adj=m.output_adjust(), has_quality=m.has_quality())
    m.close()
    return y, dc, info


def oversampled_vectors(tmpdir):
    """The host runs at 96 kHz: OversampleNAMConfig (NeuralModel.cpp:92-130) doubles every dilation and the reference leaves
    its static architectures for the dynamic path.  Same synthetic weights as the 48 kHz cases, `external_sample_rate` stored."""
    made = []
    cases = synthetic_cases()
    for j, name in enumerate(["a1_standard", "a1_nano"]):   # (A2 at 96 kHz: the reference throws json out_of_range "head_bias", it has no dynamic A2)
        case = cases[name]
        path = os.path.join(tmpdir, name + ".nam")
        with open(path, "w") as f:
            f.write(nam_text(case))
        rng = np.random.default_rng(999 + j)
        x = rng.uniform(-1.0, 1.0, N).astype(np.float32)
        try:
            y, dc, info = run_ref(path, x, external_sample_rate=96000)
        except Exception as e:   # the reference has no Internal path for this combination
            print("reference refuses", name, "at 96 kHz:", e)
            os.remove(path)
            continue
        meta = {k: v for k, v in case.items() if k != "weights"}
        out = os.path.join(HERE, "syn_%s_sr96000.npz" % name)
        np.savez_compressed(out, x=x, y=y, dc=dc, weights=np.asarray(case["weights"], dtype=np.float32), model=json.dumps(meta),
                            info=json.dumps(info), external_sample_rate=np.int32(96000))
        made.append(out)
        os.remove(path)
    return made


def main():
    O.build()
    tmpdir = os.path.join(HERE, "_tmp")
    os.makedirs(tmpdir, exist_ok=True)
    if "--a2-oversampled-only" in sys.argv:
        for m in a2_oversampled_vectors(tmpdir):
            z = np.load(m)
            print("%-48s |y|max=%.4f std=%.4f dc=%.6g info=%s" % (os.path.basename(m), np.abs(z["y"]).max(), z["y"].std(), z["dc"][-1], str(z["info"])))
        os.rmdir(tmpdir)
        return
    if "--oversampled-only" in sys.argv:
        for m in oversampled_vectors(tmpdir):
            z = np.load(m)
            print("%-48s |y|max=%.4f std=%.4f dc=%.6g info=%s" % (os.path.basename(m), np.abs(z["y"]).max(), z["y"].std(), z["dc"][-1], str(z["info"])))
        os.rmdir(tmpdir)
        return
    made = []
    fixtures = [] if ("--wide-only" in sys.argv or "--dynamic-only" in sys.argv or "--dynamic-lstm-only" in sys.argv or "--extra-lstm-only" in sys.argv) else [("BossWN-nano.nam", 1.0), ("BossWN-feather.nam", 1.0), ("BossWN-standard.nam", 1.0), ("BossWN-a2.nam", 1.0),
                ("BossWN-a2.nam", 0.0), ("BossLSTM-1x16.nam", 1.0), ("BossLSTM-2x8.nam", 1.0),
                ("tw40_blues_deluxe_deerinkstudios.json", 1.0), ("namcore_wavenet.nam", 1.0), ("namcore_lstm.nam", 1.0),
                ("namcore_wavenet_a1_standard.nam", 1.0)]
    for i, (name, q) in enumerate(fixtures):
        p = O.model_path(name)
        if p is None:
            print("missing fixture", name)
            continue
        rng = np.random.default_rng(1234 + i)
        x = rng.uniform(-1.0, 1.0, N).astype(np.float32)
        y, dc, info = run_ref(p, x, q)
        tag = os.path.splitext(name)[0].replace("-", "_") + ("" if q == 1.0 else "_q%g" % q)
        out = os.path.join(HERE, "ref_%s.npz" % tag)
        np.savez_compressed(out, x=x, y=y, dc=dc, fixture=name, quality=np.float32(q), info=json.dumps(info))
        made.append(out)
    only_dynamic = "--dynamic-only" in sys.argv   # adds the run-time-shaped cases without touching the earlier vectors
    only_wide = "--wide-only" in sys.argv         # the cases beyond two layer arrays / 32 channels, added last
    wide_only = ("dyn_3arrays", "dyn_4arrays", "dyn_48x24", "dyn_single40_k3")
    only_dyn_lstm = "--dynamic-lstm-only" in sys.argv
    only_extra = "--extra-lstm-only" in sys.argv     # the two static LSTM sizes added last (lstm_1x8, lstm_2x16)
    extra_only = ("lstm_1x8", "lstm_2x16")
    if only_dynamic or only_dyn_lstm or only_extra or only_wide:
        made = []
    allcases = list(synthetic_cases().items()) + list(dynamic_cases().items()) + list(dynamic_lstm_cases().items())
    for j, (name, case) in enumerate(allcases):
        if only_dynamic and not name.startswith("dyn_"):
            continue
        if only_dyn_lstm and not name.startswith("dyn_lstm"):
            continue
        if only_extra and name not in extra_only:
            continue
        if only_wide and name not in wide_only:
            continue
        path = os.path.join(tmpdir, name + ".nam")
        with open(path, "w") as f:
            f.write(nam_text(case))
        rng = np.random.default_rng(777 + j)
        amp = 0.5 if "lstm" in name else 1.0
        x = (rng.uniform(-1.0, 1.0, N) * amp).astype(np.float32)
        y, dc, info = run_ref(path, x)
        meta = {k: v for k, v in case.items() if k != "weights"}
        out = os.path.join(HERE, "syn_%s.npz" % name)
        np.savez_compressed(out, x=x, y=y, dc=dc, weights=np.asarray(case["weights"], dtype=np.float32), model=json.dumps(meta),
                            info=json.dumps(info))
        made.append(out)
        os.remove(path)
    os.rmdir(tmpdir)
    for m in made:
        z = np.load(m)
        print("%-48s |y|max=%.4f std=%.4f dc=%.6g size=%d" % (os.path.basename(m), np.abs(z["y"]).max(), z["y"].std(), z["dc"][-1], os.path.getsize(m)))


if __name__ == "__main__":
    main()
