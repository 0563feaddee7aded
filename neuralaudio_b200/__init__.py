"""neuralaudio_b200 -- host-side mirror of the reference's binding layer over the B200 C ABI.

The classes here play the role of ``NeuralAudioCSharp/NeuralAudio/NeuralModel.cs`` (reference :6-86): thin
objects over the C ABI of ``libneuralaudio_b200.so`` (declared in ``include/NeuralAudioCApi.h``), same names and
argument meaning as ``NeuralAudio::NeuralModelLoader`` / ``NeuralAudio::NeuralModel`` (reference
NeuralAudio/NeuralModel.h:33-231), plus the additive batched calls.

There is NO CPU fallback: importing works anywhere (so symbols can be inspected), but creating a model without the
compiled CUDA library or without a B200-class GPU raises ``NeuralAudioError``.
"""
import ctypes
import os

__all__ = ["NeuralModelLoader", "NeuralModel", "NeuralAudioError", "EModelLoadMode", "library_path", "load_library",
           "build_library", "STREAM_MAJOR", "FRAME_MAJOR", "set_option", "describe_model_file", "device_count"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libneuralaudio_b200.so"

STREAM_MAJOR = 0   # buffer[s * frames + f]
FRAME_MAJOR = 1    # buffer[f * streams + s]


class NeuralAudioError(RuntimeError):
    pass


class EModelLoadMode:   # reference NeuralModel.h:20-25
    Internal = 0
    RTNeural = 1
    NAMCore = 2


def library_path():
    # NAB200_LIBNAME: load another build of the library from the package directory (A/B timing of kernel variants)
    return os.path.join(_HERE, os.environ.get("NAB200_LIBNAME", _LIB_NAME))


def build_library(force=False):
    """Compile csrc/ for sm_100a with nvcc (no GPU needed to build)."""
    import subprocess
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", csrc, "-j8"], stdout=subprocess.DEVNULL)
    return library_path()


_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NeuralAudioError("%s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
                               "`make -C neuralaudio_b200/csrc`); neuralaudio_b200 has no CPU fallback" % path)
    L = ctypes.CDLL(path)
    vp, ci, cf, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    sig = {
        # the reference's 15 exports (NeuralAudioCApi.h:18-46)
        "CreateLoader": (vp, []),
        "DeleteLoader": (None, [vp]),
        "CreateModelFromFile": (vp, [vp, ctypes.c_wchar_p]),
        "DeleteModel": (None, [vp]),
        "SetLSTMLoadMode": (None, [vp, ci]),
        "SetWaveNetLoadMode": (None, [vp, ci]),
        "SetAudioInputLevelDBu": (None, [vp, cf]),
        "SetDefaultMaxAudioBufferSize": (None, [vp, ci]),
        "GetLoadMode": (ci, [vp]),
        "IsStatic": (ctypes.c_bool, [vp]),
        "SetMaxAudioBufferSize": (None, [vp, ci]),
        "GetRecommendedInputDBAdjustment": (cf, [vp]),
        "GetRecommendedOutputDBAdjustment": (cf, [vp]),
        "GetSampleRate": (cf, [vp]),
        "Process": (None, [vp, vp, vp, sz]),
        # additive
        "NA_GetLastError": (ctypes.c_char_p, []),
        "NA_GetVersion": (ctypes.c_char_p, []),
        "NA_GetDeviceCount": (ci, []),
        "NA_SetLoaderDevice": (None, [vp, ci]),
        "NA_SetDefaultNumStreams": (None, [vp, sz]),
        "NA_SetDefaultQualityScaleFactor": (None, [vp, cf]),
        "NA_SetExternalSampleRate": (None, [vp, ci]),
        "NA_SetCompositeModelLoadMode": (None, [vp, ci]),
        "NA_CreateModelFromMemory": (vp, [vp, ctypes.c_char_p, sz, ctypes.c_char_p, ci]),
        "NA_CreateModelFromFileEx": (vp, [vp, ctypes.c_wchar_p, ci]),
        "NA_Prewarm": (None, [vp]),
        "NA_ResetStreams": (ci, [vp]),
        "NA_HasQualityScaling": (ci, [vp]),
        "NA_GetQualityScaleFactor": (cf, [vp]),
        "NA_SetQualityScaleFactor": (None, [vp, cf]),
        "NA_GetReceptiveFieldSize": (ci, [vp]),
        "NA_GetModelVersion": (ci, [vp, ctypes.c_char_p, ci]),
        "NA_GetMetadata": (ci, [vp, ctypes.c_char_p, ctypes.c_char_p, ci]),
        "NA_SetNumStreams": (ci, [vp, sz]),
        "NA_GetNumStreams": (sz, [vp]),
        "NA_GetStateBytesPerStream": (sz, [vp]),
        "NA_GetKernelLaunchCount": (ctypes.c_ulonglong, [vp]),
        "NA_GetDevice": (ci, [vp]),
        "NA_ProcessBatch": (ci, [vp, vp, vp, sz, sz, ci]),
        "NA_ProcessBatchAsync": (ci, [vp, vp, vp, sz, sz, ci]),
        "NA_WaitBatches": (ci, [vp, ci]),
        "NA_Synchronize": (ci, [vp]),
        "NA_GetCudaStream": (vp, [vp]),
        "NA_GetDeviceBlob": (ci, [vp, ctypes.POINTER(vp), ctypes.POINTER(sz)]),
        "NA_CopyStreamState": (ci, [vp, sz, _f32p, sz]),
        "NA_DescribeModelFile": (ci, [ctypes.c_wchar_p, ci, ctypes.c_char_p, ci]),
        "NA_SetOption": (ci, [ctypes.c_char_p, ci]),
        "NA_SetLoaderOption": (None, [vp, ctypes.c_char_p, ci]),
        # multi-GPU load inside the library (NCCL through dlopen)
        "NA_NcclGetUniqueId": (ci, [ctypes.c_char_p]),
        "NA_NcclCommInitRank": (vp, [ci, ci, ctypes.c_char_p, ci]),
        "NA_NcclCommCount": (ci, [vp]),
        "NA_NcclCommDestroy": (None, [vp]),
        "NA_BroadcastModel": (ctypes.c_longlong, [vp, vp, ci]),
        "NA_BroadcastModelOnComm": (ctypes.c_longlong, [vp, vp, ci]),
        "NA_CreateModelSharded": (vp, [vp, ctypes.c_wchar_p, ctypes.POINTER(ci), ci, ci]),
        "NA_GetNumShards": (ci, [vp]),
        "NA_GetBroadcastBytes": (ctypes.c_longlong, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError here == a declared symbol is missing from the build
        fn.restype = res
        fn.argtypes = args
    L._na_signatures = sig
    _lib = L
    return L


def _last_error(L):
    msg = L.NA_GetLastError()
    return msg.decode("utf-8", "replace") if msg else ""


def device_count():
    return load_library().NA_GetDeviceCount()


def _prefer_bundled_nccl():
    """The library reaches NCCL with dlopen("libnccl.so.2").  In a Python process that also imports PyTorch, the copy that
    gets loaded FIRST wins for everybody (same soname), and PyTorch needs the one bundled with it (nvidia/nccl/lib): point
    the library at that copy unless the user chose one (NAB200_NCCL_LIB)."""
    if os.environ.get("NAB200_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            for base in spec.submodule_search_locations:
                cand = os.path.join(base, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["NAB200_NCCL_LIB"] = cand
                    return
    except Exception:
        pass


def nccl_get_unique_id():
    """128 bytes from ncclGetUniqueId (root rank); ship them to the other ranks by any means, then NcclComm(...) everywhere."""
    _prefer_bundled_nccl()
    L = load_library()
    buf = ctypes.create_string_buffer(128)
    if L.NA_NcclGetUniqueId(buf) != 0:
        raise NeuralAudioError(_last_error(L) or "NCCL unavailable")
    return buf.raw


class NcclComm:
    """One rank's NCCL communicator owned by the library (ncclCommInitRank), for NeuralModel.BroadcastModel."""

    def __init__(self, nranks, rank, unique_id, device):
        _prefer_bundled_nccl()
        self._L = load_library()
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of nccl_get_unique_id()")
        self._h = self._L.NA_NcclCommInitRank(int(nranks), int(rank), bytes(unique_id), int(device))
        if not self._h:
            raise NeuralAudioError(_last_error(self._L) or "ncclCommInitRank failed")
        self.nranks = self._L.NA_NcclCommCount(self._h)
        self.rank = int(rank)

    def close(self):
        if self._h:
            self._L.NA_NcclCommDestroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_option(name, value):
    return load_library().NA_SetOption(name.encode(), int(value))


def describe_model_file(path, external_sample_rate=48000):
    """Host-only parse/dispatch/pack report (JSON text -> dict); touches no GPU."""
    import json
    L = load_library()
    buf = ctypes.create_string_buffer(1 << 16)
    n = L.NA_DescribeModelFile(os.path.abspath(path), int(external_sample_rate), buf, len(buf))
    if n < 0:
        raise NeuralAudioError(_last_error(L))
    return json.loads(buf.value.decode())


def _pointer_of(x, writable=False):
    """(address, element count, keepalive) of a numpy array or torch tensor holding contiguous float32."""
    if hasattr(x, "data_ptr"):   # torch tensor (CPU, pinned or CUDA)
        import torch
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise NeuralAudioError("buffers must be contiguous float32")
        return x.data_ptr(), x.numel(), x
    import numpy as np
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
        raise NeuralAudioError("buffers must be contiguous float32 numpy arrays or torch tensors")
    if writable and not x.flags["WRITEABLE"]:
        raise NeuralAudioError("output buffer is read-only")
    return x.ctypes.data, x.size, x


class NeuralModelLoader:
    """Mirror of NeuralAudio::NeuralModelLoader (reference NeuralModel.h:148-231)."""

    def __init__(self):
        self._L = load_library()
        self._h = self._L.CreateLoader()

    def close(self):
        if self._h:
            self._L.DeleteLoader(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def SetLSTMLoadMode(self, mode):
        self._L.SetLSTMLoadMode(self._h, int(mode))

    def SetWaveNetLoadMode(self, mode):
        self._L.SetWaveNetLoadMode(self._h, int(mode))

    def SetAudioInputLevelDBu(self, dbu):
        self._L.SetAudioInputLevelDBu(self._h, float(dbu))

    def SetDefaultMaxAudioBufferSize(self, n):
        self._L.SetDefaultMaxAudioBufferSize(self._h, int(n))

    def SetDefaultQualityScaleFactor(self, q):
        self._L.NA_SetDefaultQualityScaleFactor(self._h, float(q))

    def SetExternalSampleRate(self, sr):
        self._L.NA_SetExternalSampleRate(self._h, int(sr))

    def SetCompositeModelLoadMode(self, mode):
        self._L.NA_SetCompositeModelLoadMode(self._h, int(mode))

    def SetOption(self, name, value):
        """Tuning knob for the models this loader builds (the process-wide set_option defaults stay untouched)."""
        self._L.NA_SetLoaderOption(self._h, name.encode(), int(value))

    def SetDevice(self, device):
        self._L.NA_SetLoaderDevice(self._h, int(device))

    def SetDefaultNumStreams(self, n):
        self._L.NA_SetDefaultNumStreams(self._h, int(n))

    def CreateFromFile(self, path, doPrewarm=True):
        h = self._L.NA_CreateModelFromFileEx(self._h, os.path.abspath(path), 1 if doPrewarm else 0)
        if not h:
            raise NeuralAudioError(_last_error(self._L) or "model could not be loaded")
        return NeuralModel(self._L, h)

    def CreateShardedFromFile(self, path, devices, doPrewarm=True):
        """One process, several GPUs: the model on every listed device, one grouped ncclBroadcast of [weights | prewarmed state]
        from the first, the stream batch (SetDefaultNumStreams = the total) cut into contiguous shards."""
        _prefer_bundled_nccl()
        arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = self._L.NA_CreateModelSharded(self._h, os.path.abspath(path), arr, len(devices), 1 if doPrewarm else 0)
        if not h:
            raise NeuralAudioError(_last_error(self._L) or "model could not be loaded")
        return NeuralModel(self._L, h)

    def CreateFromMemory(self, data, extension=".nam", doPrewarm=True):
        if isinstance(data, str):
            data = data.encode()
        h = self._L.NA_CreateModelFromMemory(self._h, data, len(data), extension.encode(), 1 if doPrewarm else 0)
        if not h:
            raise NeuralAudioError(_last_error(self._L) or "model could not be loaded")
        return NeuralModel(self._L, h)


class NeuralModel:
    """Mirror of NeuralAudio::NeuralModel (reference NeuralModel.h:33-146) plus the additive stream-slot API."""

    def __init__(self, L, handle):
        self._L = L
        self._h = handle

    def close(self):
        if self._h:
            self._L.DeleteModel(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- reference surface
    def GetLoadMode(self):
        return self._L.GetLoadMode(self._h)

    def IsStatic(self):
        return bool(self._L.IsStatic(self._h))

    def SetMaxAudioBufferSize(self, n):
        self._L.SetMaxAudioBufferSize(self._h, int(n))

    def GetRecommendedInputDBAdjustment(self):
        return self._L.GetRecommendedInputDBAdjustment(self._h)

    def GetRecommendedOutputDBAdjustment(self):
        return self._L.GetRecommendedOutputDBAdjustment(self._h)

    def GetSampleRate(self):
        return self._L.GetSampleRate(self._h)

    def GetReceptiveFieldSize(self):
        return self._L.NA_GetReceptiveFieldSize(self._h)

    def HasQualityScaling(self):
        return bool(self._L.NA_HasQualityScaling(self._h))

    def GetQualityScaleFactor(self):
        return self._L.NA_GetQualityScaleFactor(self._h)

    def SetQualityScaleFactor(self, q):
        self._L.NA_SetQualityScaleFactor(self._h, float(q))

    def GetModelVersion(self):
        buf = ctypes.create_string_buffer(256)
        self._L.NA_GetModelVersion(self._h, buf, len(buf))
        return buf.value.decode()

    def GetMetadata(self, key):
        buf = ctypes.create_string_buffer(1 << 16)
        self._L.NA_GetMetadata(self._h, key.encode(), buf, len(buf))
        return buf.value.decode()

    def Prewarm(self):
        self._L.NA_Prewarm(self._h)

    def Process(self, input, output=None):
        """One mono stream (slot 0).  numpy in -> numpy out (synchronous, like the reference)."""
        import numpy as np
        if output is None:
            output = np.empty_like(input) if isinstance(input, np.ndarray) else input.new_empty(input.shape)
        pi, ni, _k1 = _pointer_of(input)
        po, no, _k2 = _pointer_of(output, True)
        if ni != no:
            raise NeuralAudioError("input/output size mismatch")
        self._L.Process(self._h, pi, po, ni)
        err = _last_error(self._L)
        if err:
            raise NeuralAudioError(err)
        return output

    # --- additive
    def SetNumStreams(self, n):
        if self._L.NA_SetNumStreams(self._h, int(n)) != 0:
            raise NeuralAudioError(_last_error(self._L))

    def GetNumStreams(self):
        return self._L.NA_GetNumStreams(self._h)

    def BroadcastModel(self, comm, root=0):
        """ONE ncclBroadcast per resident engine of [packed weights | prewarmed state template] from `root`; returns the bytes."""
        n = self._L.NA_BroadcastModel(self._h, comm._h, int(root))
        if n < 0:
            raise NeuralAudioError(_last_error(self._L) or "BroadcastModel failed")
        return int(n)

    def GetNumShards(self):
        return int(self._L.NA_GetNumShards(self._h))

    def GetBroadcastBytes(self):
        return int(self._L.NA_GetBroadcastBytes(self._h))

    def GetKernelLaunchCount(self):
        return int(self._L.NA_GetKernelLaunchCount(self._h))

    def GetStateBytesPerStream(self):
        return self._L.NA_GetStateBytesPerStream(self._h)

    def GetDevice(self):
        return self._L.NA_GetDevice(self._h)

    def ResetStreams(self):
        if self._L.NA_ResetStreams(self._h) != 0:
            raise NeuralAudioError(_last_error(self._L))

    def ProcessBatch(self, input, output, numStreams, numFrames, layout=STREAM_MAJOR):
        pi, ni, _k1 = _pointer_of(input)
        po, no, _k2 = _pointer_of(output, True)
        if ni < numStreams * numFrames or no < numStreams * numFrames:
            raise NeuralAudioError("buffer smaller than numStreams * numFrames")
        if self._L.NA_ProcessBatch(self._h, pi, po, int(numStreams), int(numFrames), int(layout)) != 0:
            raise NeuralAudioError(_last_error(self._L))
        return output

    def ProcessBatchAsync(self, input, output, numStreams, numFrames, layout=STREAM_MAJOR):
        """Pipelined ProcessBatch for page-locked host buffers (or device buffers): returns once queued."""
        pi, ni, _k1 = _pointer_of(input)
        po, no, _k2 = _pointer_of(output, True)
        if ni < numStreams * numFrames or no < numStreams * numFrames:
            raise NeuralAudioError("buffer smaller than numStreams * numFrames")
        if self._L.NA_ProcessBatchAsync(self._h, pi, po, int(numStreams), int(numFrames), int(layout)) != 0:
            raise NeuralAudioError(_last_error(self._L))
        return output

    def WaitBatches(self, lag=0):
        if self._L.NA_WaitBatches(self._h, int(lag)) != 0:
            raise NeuralAudioError(_last_error(self._L))

    def Synchronize(self):
        if self._L.NA_Synchronize(self._h) != 0:
            raise NeuralAudioError(_last_error(self._L))

    def GetCudaStream(self):
        return self._L.NA_GetCudaStream(self._h)

    def GetDeviceBlob(self):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        if self._L.NA_GetDeviceBlob(self._h, ctypes.byref(p), ctypes.byref(n)) != 0:
            raise NeuralAudioError(_last_error(self._L))
        return p.value, n.value

    def CopyStreamState(self, stream=0):
        import numpy as np
        n = self.GetStateBytesPerStream() // 4
        out = np.empty(n, dtype=np.float32)
        got = self._L.NA_CopyStreamState(self._h, int(stream), out.ctypes.data_as(_f32p), n)
        if got < 0:
            raise NeuralAudioError(_last_error(self._L))
        return out[:got]
