"""ctypes binding of the UNMODIFIED reference codec built by oracle/Makefile into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Nothing under solo_b200/ does.

API bound: JC1_SDK_SRC_{ARM,FLP}/interface/AGR_JC1_SDK_API.h:33-64 (six functions).
Internal reference functions (exported by the .so because they are non-static C) are reachable
through ``RefLib.lib`` for unit-level parity tests of individual primitives.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class EncCtrl(C.Structure):
    # AGR_JC1_SDK_API.h:11-21
    _fields_ = [(n, C.c_int32) for n in ("mode", "targetRate_bps", "samplerate", "dtx_enable",
                                         "framesize_ms", "joint_enable", "joint_mode", "useMDIndex")]


class DecCtrl(C.Structure):
    # AGR_JC1_SDK_API.h:23-31
    _fields_ = [(n, C.c_int32) for n in ("packetLoss_perc", "samplerate", "framesize_ms",
                                         "joint_enable", "joint_mode", "useMDIndex")]


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libjc1_fix.so", "libjc1_flp.so"))


class RefLib:
    def __init__(self, kind="fix"):
        path = os.path.join(REF_DIR, "libjc1_%s.so" % kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run `make -C oracle` (needs /root/reference)")
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        L.AGR_Sate_Encoder_Init.restype = C.c_void_p
        L.AGR_Sate_Encoder_Init.argtypes = [C.POINTER(EncCtrl)]
        L.AGR_Sate_Encoder_Encode.restype = C.c_int32
        L.AGR_Sate_Encoder_Encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.AGR_Sate_Encoder_Uninit.argtypes = [C.c_void_p]
        L.AGR_Sate_Decoder_Init.restype = C.c_void_p
        L.AGR_Sate_Decoder_Init.argtypes = [C.POINTER(DecCtrl)]
        L.AGR_Sate_Decoder_Decode.restype = C.c_int32
        L.AGR_Sate_Decoder_Decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.AGR_Sate_Decoder_Uninit.argtypes = [C.c_void_p]


_libs = {}


def lib(kind):
    if kind not in _libs:
        _libs[kind] = RefLib(kind)
    return _libs[kind]


class RefEncoder:
    """One reference encoder handle (kind = 'fix' | 'flp')."""

    def __init__(self, kind="fix", rate=13600, dtx=0, use_md_index=0, framesize_ms=40, joint_hb=0):
        self.L = lib(kind).lib
        self.samples = 16 * framesize_ms
        self.ctrl = EncCtrl(2, rate, 16000, dtx, framesize_ms, 1 if joint_hb else 0, 1 if joint_hb else 0, use_md_index)
        self.h = self.L.AGR_Sate_Encoder_Init(C.byref(self.ctrl))
        self._bits = (C.c_uint8 * 1024)()
        self._nb = (C.c_int16 * 6)()

    def encode(self, pcm640):
        pcm = np.ascontiguousarray(pcm640, dtype=np.int16)
        assert pcm.size == self.samples
        C.memset(self._nb, 0, 12)
        n = self.L.AGR_Sate_Encoder_Encode(self.h, pcm.ctypes.data, self._bits, 1024, self._nb)
        return bytes(self._bits[:max(n, 0)]), (int(self._nb[0]), int(self._nb[1])), n

    def close(self):
        if self.h:
            self.L.AGR_Sate_Encoder_Uninit(self.h)
            self.h = None

    __del__ = close


class RefDecoder:
    """One reference decoder handle (kind = 'flp' is the PCM parity target)."""

    def __init__(self, kind="flp", use_md_index=0, framesize_ms=40, joint_hb=0):
        self.L = lib(kind).lib
        self.samples = 16 * framesize_ms
        self.ctrl = DecCtrl(0, 16000, framesize_ms, 1 if joint_hb else 0, 1 if joint_hb else 0, use_md_index)
        self.h = self.L.AGR_Sate_Decoder_Init(C.byref(self.ctrl))
        self._out = np.zeros(960, dtype=np.int16)
        self._ns = C.c_int16(0)

    def decode(self, payload, nbytes, lostflag):
        """payload/nbytes pre-trimmed by the caller as dec_main.c:245-307 does."""
        buf = (C.c_uint8 * 1024)(*payload) if payload else (C.c_uint8 * 1024)()
        nb = (C.c_int16 * 6)(int(nbytes[0]), int(nbytes[1]))
        ret = self.L.AGR_Sate_Decoder_Decode(self.h, self._out.ctypes.data, C.byref(self._ns), buf, nb, int(lostflag))
        return self._out[:self.samples].copy(), ret

    def close(self):
        if self.h:
            self.L.AGR_Sate_Decoder_Uninit(self.h)
            self.h = None

    __del__ = close
