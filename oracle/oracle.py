"""TEST INFRASTRUCTURE ONLY -- Python front-end of the oracle.

Two checkers live here:

* ``PortModel``  -- the plain-C restatement in ``na_oracle.c`` (kind "port").  The loader dispatch below restates
  ``NeuralModelLoader::CreateFromJson`` (reference NeuralAudio/NeuralModel.cpp:338-581) for the Internal branch.
* ``RefModel``   -- the UNMODIFIED reference compiled by ``build_ref.sh`` into ``oracle/_ref/libna_ref*.so``
  (kind "reference"), driven through its own C ABI (NeuralAudioCAPI/NeuralAudioCApi.h:18-46) exactly like the
  C# binding does (NeuralAudioCSharp/NeuralAudio/NativeApi.cs:11-54).

The product (neuralaudio_b200/) must never import this module.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
MODELS_DIR = os.path.join(REF_DIR, "models")

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the C restatement (and the reference, when /root/reference exists)."""
    so = os.path.join(HERE, "libna_oracle.so")
    src = os.path.join(HERE, "na_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libna_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class _ArrayDesc(ctypes.Structure):
    _fields_ = [("input_size", ctypes.c_int), ("channels", ctypes.c_int), ("head_size", ctypes.c_int),
                ("head_kernel", ctypes.c_int), ("head_bias", ctypes.c_int), ("num_layers", ctypes.c_int),
                ("activation", ctypes.c_int), ("kernel_sizes", _c_int_p), ("dilations", _c_int_p), ("head_dilation", ctypes.c_int)]


_port_lib = None


def _port():
    global _port_lib
    if _port_lib is None:
        so = os.path.join(HERE, "libna_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.na_oracle_wavenet_create.restype = ctypes.c_void_p
        L.na_oracle_wavenet_create.argtypes = [ctypes.c_int, ctypes.POINTER(_ArrayDesc), _c_float_p, ctypes.c_int]
        L.na_oracle_wavenet_prewarm.argtypes = [ctypes.c_void_p]
        L.na_oracle_wavenet_process.argtypes = [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_int]
        L.na_oracle_wavenet_receptive_field.argtypes = [ctypes.c_void_p]
        L.na_oracle_wavenet_receptive_field.restype = ctypes.c_int
        L.na_oracle_wavenet_destroy.argtypes = [ctypes.c_void_p]
        L.na_oracle_lstm_create_nam.restype = ctypes.c_void_p
        L.na_oracle_lstm_create_nam.argtypes = [ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_int]
        L.na_oracle_lstm_create_keras.restype = ctypes.c_void_p
        L.na_oracle_lstm_create_keras.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_c_float_p),
                                                  ctypes.POINTER(_c_float_p), ctypes.POINTER(_c_float_p),
                                                  _c_float_p, ctypes.c_float]
        L.na_oracle_lstm_process.argtypes = [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_int]
        L.na_oracle_lstm_prewarm.argtypes = [ctypes.c_void_p]
        L.na_oracle_lstm_destroy.argtypes = [ctypes.c_void_p]
        _port_lib = L
    return _port_lib


# NeuralModel.cpp:71-76 -- the official dilation / kernel tables
STD_DILATIONS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512]
LITE_DILATIONS = [1, 2, 4, 8, 16, 32, 64]
LITE_DILATIONS2 = [128, 256, 512, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512]
A2_KERNEL_SIZES = [6] * 14 + [15, 15] + [6] * 7
A2_DILATIONS = [1, 3, 7, 17, 41, 101, 239] * 2 + [1, 13] + [1, 3, 7, 17, 41, 101, 239]


def nam_is_a2(version):
    """NAMIsA2, NeuralModel.cpp:159-168"""
    parts = (version.split(".") + ["0", "0", "0"])[:3]
    major, minor, patch = (int(p) for p in parts)
    return major > 0 or minor > 5 or (minor == 5 and patch > 4)


def oversample_nam_config(model_json, external_sample_rate):
    """OversampleNAMConfig, NeuralModel.cpp:92-130 (mutates model_json)"""
    if model_json["architecture"] != "WaveNet":
        return
    sr = 48000
    if isinstance(model_json.get("sample_rate"), (int, float)):
        sr = int(model_json["sample_rate"])
    if sr == external_sample_rate or external_sample_rate % sr != 0:
        return
    f = external_sample_rate // sr
    for layer in model_json["config"]["layers"]:
        layer["dilations"] = [int(d) * f for d in layer["dilations"]]
        if "head" in layer:
            layer["head"]["head_dilation"] = f


def wavenet_arrays_from_config(config, version):
    """Architecture descriptors for the Internal WaveNet branch (NeuralModel.cpp:383-478).

    Returns a list of dicts.  A1-style configs (``kernel_size``/``head_size``) map to Tanh + 1x1 head
    (InternalModel.h:152-159, WaveNetDynamic.h); the single-array A2 layout maps to LeakyReLU + K=16 head
    (NeuralModel.cpp:389-421)."""
    layers = config["layers"]
    arrays = []
    if len(layers) == 1 and "kernel_sizes" in layers[0]:
        lc = layers[0]
        if lc["channels"] not in (3, 8) or len(lc["dilations"]) != len(lc["kernel_sizes"]):
            raise ValueError("A2 WaveNet outside the Internal static set (reference would use NAM Core)")
        # (standard dilations: the reference's static A2; any other list - e.g. doubled by OversampleNAMConfig together with the
        # head dilation - is the same network with other delays: the reference runs it on NAM Core, the port follows its arithmetic)
        arrays.append(dict(input_size=1, channels=int(lc["channels"]), head_size=1, head_kernel=int(lc["head"].get("kernel_size", 16)), head_bias=1,
                           kernel_sizes=[int(k) for k in lc["kernel_sizes"]], dilations=[int(d) for d in lc["dilations"]], activation=1,
                           head_dilation=int(lc["head"].get("head_dilation", 1))))
        return arrays
    for lc in layers:
        n = len(lc["dilations"])
        arrays.append(dict(input_size=int(lc["input_size"]), channels=int(lc["channels"]),
                           head_size=int(lc["head_size"]), head_kernel=1, head_bias=1 if lc["head_bias"] else 0,
                           kernel_sizes=[int(lc["kernel_size"])] * n, dilations=[int(d) for d in lc["dilations"]],
                           activation=0))
    return arrays


class PortModel:
    """The plain-C restatement behind a NeuralModel-like surface (Process / Prewarm / quality)."""

    def __init__(self, model_json, ext=".nam", prewarm=True, quality=1.0, external_sample_rate=48000):
        self._L = _port()
        self._subs = []      # [(max_value, handle, kind)]
        self._cur = 0
        self._keep = []
        if ext == ".nam":
            model_json = json.loads(json.dumps(model_json))
            if model_json["architecture"] == "SlimmableContainer":
                # ScalableCompositeModel::CreateModelFromNAMJson, CompositeModel.h:137-159
                levels = []
                for i, sm in enumerate(model_json["config"]["submodels"]):
                    oversample_nam_config(sm["model"], external_sample_rate)
                    self._subs.append(self._create_nam(sm["model"]))
                    levels.append((float(sm["max_value"]), i))
                self._levels = sorted(levels, key=lambda t: t[0])
                self.set_quality(quality)
            else:
                oversample_nam_config(model_json, external_sample_rate)
                self._subs.append(self._create_nam(model_json))
                self._levels = [(1.0, 0)]
        else:
            self._subs.append(self._create_keras(model_json))
            self._levels = [(1.0, 0)]
        if prewarm:
            self.prewarm()

    def _create_nam(self, mj):
        w = np.ascontiguousarray(np.asarray(mj["weights"], dtype=np.float32))
        arch = mj["architecture"]
        if arch == "WaveNet":
            arrays = wavenet_arrays_from_config(mj["config"], mj["version"])
            descs = (_ArrayDesc * len(arrays))()
            for i, a in enumerate(arrays):
                ks = (ctypes.c_int * len(a["kernel_sizes"]))(*a["kernel_sizes"])
                ds = (ctypes.c_int * len(a["dilations"]))(*a["dilations"])
                self._keep += [ks, ds]
                descs[i] = _ArrayDesc(a["input_size"], a["channels"], a["head_size"], a["head_kernel"], a["head_bias"],
                                      len(a["dilations"]), a["activation"], ks, ds, a.get("head_dilation", 1))
            h = self._L.na_oracle_wavenet_create(len(arrays), descs, w.ctypes.data_as(_c_float_p), int(w.size))
            if not h:
                raise RuntimeError("Wrong number of weights")   # WaveNet.h:704-709
            return (h, "wavenet")
        if arch == "LSTM":
            c = mj["config"]
            h = self._L.na_oracle_lstm_create_nam(int(c["num_layers"]), int(c["hidden_size"]),
                                                  w.ctypes.data_as(_c_float_p), int(w.size))
            if not h:
                raise RuntimeError("Wrong number of weights")
            return (h, "lstm")
        raise ValueError("unsupported architecture " + arch)

    def _create_keras(self, mj):
        # InternalLSTMModelT::CreateModelFromKerasJson, InternalModel.h:297-356
        layers = mj["layers"]
        if len(layers) < 2 or layers[-1]["type"] != "dense":
            raise ValueError("unsupported keras model")
        H = int(layers[0]["shape"][-1])
        L = len(layers) - 1
        flat = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))
        ker, rec, bia = [], [], []
        for l in layers[:-1]:
            if l["type"] != "lstm":
                raise ValueError("unsupported keras layer " + l["type"])
            ker.append(flat(l["weights"][0])); rec.append(flat(l["weights"][1])); bia.append(flat(l["weights"][2]))
        hw = flat(layers[-1]["weights"][0])
        hb = float(layers[-1]["weights"][1][0])
        self._keep += ker + rec + bia + [hw]
        P = lambda arrs: (_c_float_p * len(arrs))(*[a.ctypes.data_as(_c_float_p) for a in arrs])
        h = self._L.na_oracle_lstm_create_keras(L, H, P(ker), P(rec), P(bia), hw.ctypes.data_as(_c_float_p), hb)
        return (h, "lstm")

    def set_quality(self, q):
        # GetModelIndexFromQualityScale, CompositeModel.h:200-213
        idx = 0
        for level, i in self._levels:
            idx = i
            if q <= level:
                break
        self._cur = idx

    def prewarm(self):
        # CompositeModel::Prewarm in LoadAll mode prewarms every sub-model (CompositeModel.h:102-118)
        for h, kind in self._subs:
            if kind == "wavenet":
                self._L.na_oracle_wavenet_prewarm(h)
            else:
                self._L.na_oracle_lstm_prewarm(h)

    def receptive_field(self):
        h, kind = self._subs[self._cur]
        return self._L.na_oracle_wavenet_receptive_field(h) if kind == "wavenet" else -1

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        h, kind = self._subs[self._cur]
        fn = self._L.na_oracle_wavenet_process if kind == "wavenet" else self._L.na_oracle_lstm_process
        fn(h, x.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p), int(x.size))
        return out

    def close(self):
        for h, kind in self._subs:
            (self._L.na_oracle_wavenet_destroy if kind == "wavenet" else self._L.na_oracle_lstm_destroy)(h)
        self._subs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_file(cls, path, **kw):
        with open(path) as f:
            mj = json.load(f)
        return cls(mj, ext=os.path.splitext(path)[1], **kw)


# ------------------------------------------------------------------------------------------------------
# the compiled reference
# ------------------------------------------------------------------------------------------------------

_ref_lib = None


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "libna_ref.so"))


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return all(k in txt for k in ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd"))
    except OSError:
        return False


def ref_lib_path():
    # NA_REF_NAMCORE=1: the build that carries the reference's NAM Core back-end (golden vectors of the files its Internal path refuses)
    nc = os.path.join(REF_DIR, "libna_ref_namcore.so")
    if os.environ.get("NA_REF_NAMCORE", "") == "1" and os.path.exists(nc):
        return nc
    v4 = os.path.join(REF_DIR, "libna_ref_v4.so")
    if _cpu_has_avx512() and os.path.exists(v4) and os.environ.get("NA_REF_ISA", "") != "v3":
        return v4
    return os.path.join(REF_DIR, "libna_ref.so")


def _ref():
    global _ref_lib
    if _ref_lib is None:
        if not ref_available():
            build()
        L = ctypes.CDLL(ref_lib_path())
        L.CreateLoader.restype = ctypes.c_void_p
        L.DeleteLoader.argtypes = [ctypes.c_void_p]
        L.CreateModelFromFile.restype = ctypes.c_void_p
        L.CreateModelFromFile.argtypes = [ctypes.c_void_p, ctypes.c_wchar_p]
        L.RefX_CreateModelFromFileNoPrewarm.restype = ctypes.c_void_p
        L.RefX_CreateModelFromFileNoPrewarm.argtypes = [ctypes.c_void_p, ctypes.c_wchar_p]
        L.DeleteModel.argtypes = [ctypes.c_void_p]
        L.SetDefaultMaxAudioBufferSize.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.SetAudioInputLevelDBu.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.SetLSTMLoadMode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.SetWaveNetLoadMode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.GetLoadMode.argtypes = [ctypes.c_void_p]
        L.IsStatic.argtypes = [ctypes.c_void_p]
        L.IsStatic.restype = ctypes.c_bool
        L.SetMaxAudioBufferSize.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for fn in ("GetRecommendedInputDBAdjustment", "GetRecommendedOutputDBAdjustment", "GetSampleRate"):
            getattr(L, fn).argtypes = [ctypes.c_void_p]
            getattr(L, fn).restype = ctypes.c_float
        L.Process.argtypes = [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_size_t]
        L.RefX_SetQualityScaleFactor.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.RefX_GetQualityScaleFactor.argtypes = [ctypes.c_void_p]
        L.RefX_GetQualityScaleFactor.restype = ctypes.c_float
        L.RefX_HasQualityScaling.argtypes = [ctypes.c_void_p]
        L.RefX_Prewarm.argtypes = [ctypes.c_void_p]
        L.RefX_GetReceptiveFieldSize.argtypes = [ctypes.c_void_p]
        L.RefX_IsNull.argtypes = [ctypes.c_void_p]
        L.RefX_SetDefaultQualityScaleFactor.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.RefX_SetExternalSampleRate.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.RefX_GetMetadata.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.RefX_GetModelVersion.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.RefX_BenchProcess.restype = ctypes.c_double
        L.RefX_BenchProcess.argtypes = [ctypes.c_wchar_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_uint, ctypes.POINTER(ctypes.c_double)]
        L.RefX_BenchSteps.restype = ctypes.c_double
        L.RefX_BenchSteps.argtypes = [ctypes.c_wchar_p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        _ref_lib = L
    return _ref_lib


class RefModel:
    """One model object of the compiled reference (one mono stream, NeuralModel.h:127)."""

    def __init__(self, path, prewarm=True, quality=1.0, max_buffer=128, external_sample_rate=48000, input_dbu=None):
        L = self._L = _ref()
        self._loader = L.CreateLoader()
        L.SetDefaultMaxAudioBufferSize(self._loader, max_buffer)
        L.RefX_SetDefaultQualityScaleFactor(self._loader, quality)
        L.RefX_SetExternalSampleRate(self._loader, external_sample_rate)
        if input_dbu is not None:
            L.SetAudioInputLevelDBu(self._loader, input_dbu)
        create = L.CreateModelFromFile if prewarm else L.RefX_CreateModelFromFileNoPrewarm
        self._m = create(self._loader, os.path.abspath(path))
        if not self._m or L.RefX_IsNull(self._m):
            raise RuntimeError("reference failed to load " + path)

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._L.Process(self._m, x.ctypes.data_as(_c_float_p), out.ctypes.data_as(_c_float_p), x.size)
        return out

    def process_blocks(self, x, block):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        for i in range(0, x.size, block):
            out[i:i + block] = self.process(x[i:i + block])
        return out

    def set_quality(self, q):
        self._L.RefX_SetQualityScaleFactor(self._m, q)

    def prewarm(self):
        self._L.RefX_Prewarm(self._m)

    def receptive_field(self):
        return self._L.RefX_GetReceptiveFieldSize(self._m)

    def load_mode(self):
        return self._L.GetLoadMode(self._m)

    def is_static(self):
        return bool(self._L.IsStatic(self._m))

    def sample_rate(self):
        return self._L.GetSampleRate(self._m)

    def input_adjust(self):
        return self._L.GetRecommendedInputDBAdjustment(self._m)

    def output_adjust(self):
        return self._L.GetRecommendedOutputDBAdjustment(self._m)

    def has_quality(self):
        return bool(self._L.RefX_HasQualityScaling(self._m))

    def metadata(self, key):
        buf = ctypes.create_string_buffer(65536)
        self._L.RefX_GetMetadata(self._m, key.encode(), buf, 65536)
        return buf.value.decode()

    def version(self):
        buf = ctypes.create_string_buffer(256)
        self._L.RefX_GetModelVersion(self._m, buf, 256)
        return buf.value.decode()

    def close(self):
        if self._m:
            self._L.DeleteModel(self._m)
            self._L.DeleteLoader(self._loader)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_bench(path, frames, seconds, threads, instances_per_thread=64, quality=1.0, seed=1234):
    """Aggregate samples/s of the reference's own Process() on `threads` host threads
    (protocol of BASELINE.md §4 / Utils/ModelTest/ModelTest.cpp:59-79, on white noise)."""
    L = _ref()
    per = (ctypes.c_double * threads)()
    total = L.RefX_BenchProcess(os.path.abspath(path), quality, threads, instances_per_thread, frames,
                                float(seconds), seed, per)
    return total, list(per)


def ref_bench_steps(path, frames, threads, instances_per_thread, warmup, steps, quality=1.0, seed=1234):
    """Seconds for `steps` steps; one step = every one of threads*instances_per_thread reference model objects
    processes one `frames`-sample block (its own stream of white noise)."""
    return _ref().RefX_BenchSteps(os.path.abspath(path), quality, threads, instances_per_thread, frames, warmup, steps, seed)


def model_path(name):
    p = os.path.join(MODELS_DIR, name)
    return p if os.path.exists(p) else None
