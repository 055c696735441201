"""ctypes driver of tests/_hostsim/libsb_hostsim.so: the per-stream kernel source (solo_b200/csrc/*.cuh) compiled by g++.

TEST INFRASTRUCTURE ONLY.  It exists so that the CPU-only container can check the kernel arithmetic bit for bit; the
product (libsolo_b200.so) never links or loads it.  Same call shapes as oracle/ref.py so tests can swap the two."""
import ctypes as C

import numpy as np

from . import build_hostsim

_libs = {}
EMU = False   # set tests.hostsim.sim.EMU = True (or use emu=True) to drive the 32-thread cooperative build


def lib(emu=None):
    emu = EMU if emu is None else emu
    if emu not in _libs:
        L = C.CDLL(build_hostsim.build(emu=emu))
        L.hs_enc_create.restype = C.c_void_p
        L.hs_enc_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.hs_enc_create2.restype = C.c_void_p
        L.hs_enc_create2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.hs_dec_create2.restype = C.c_void_p
        L.hs_dec_create2.argtypes = [C.c_int, C.c_int]
        L.hs_enc_create3.restype = C.c_void_p
        L.hs_enc_create3.argtypes = [C.c_int] * 5
        L.hs_dec_create3.restype = C.c_void_p
        L.hs_dec_create3.argtypes = [C.c_int] * 3
        L.hs_enc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_enc_destroy.argtypes = [C.c_void_p]
        L.hs_dec_create.restype = C.c_void_p
        L.hs_dec_create.argtypes = [C.c_int]
        L.hs_dec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.hs_dec_destroy.argtypes = [C.c_void_p]
        if emu:
            L.hs_set_emu_nsq.argtypes = [C.c_int]
            L.hs_enc_encode_pair.argtypes = [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3
        _libs[emu] = L
    return _libs[emu]


class SimEncoder:
    def __init__(self, rate=13600, dtx=0, use_md_index=0, cap=1024, emu=None, framesize_ms=40, joint_hb=0, emu_nsq=False):
        """emu_nsq (emu build only): the quantiser stage runs the device kernel's code (sb_nsq_warp.cuh) on emulated lanes too."""
        self.L = lib(emu)
        self.emu_nsq = bool(emu_nsq)
        assert not self.emu_nsq or self.L.hs_is_emu() == 1
        self.h = self.L.hs_enc_create3(rate, dtx, use_md_index, framesize_ms, joint_hb)
        self.samples = 16 * framesize_ms
        self.cap = cap
        self.out = np.zeros(cap, np.uint8)
        self.nb = np.zeros(6, np.int16)

    def encode(self, pcm640):
        x = np.ascontiguousarray(pcm640, np.int16)
        assert x.size == self.samples
        if self.L.hs_is_emu() == 1:
            self.L.hs_set_emu_nsq(1 if self.emu_nsq else 0)
        n = self.L.hs_enc_encode(self.h, x.ctypes.data, self.out.ctypes.data, self.cap, self.nb.ctypes.data)
        return bytes(self.out[:max(n, 0)]), (int(self.nb[0]), int(self.nb[1])), n

    def close(self):
        if self.h:
            self.L.hs_enc_destroy(self.h)
            self.h = None


def encode_pair(ea, eb, pcm_a, pcm_b):
    """Two encoders of the emu="gw16" build through ONE emulated quantiser warp (a lane group each).  Returns the two
    (payload, (n0, n1), n) results."""
    xa, xb = np.ascontiguousarray(pcm_a, np.int16), np.ascontiguousarray(pcm_b, np.int16)
    ret = np.zeros(2, np.int32)
    r = ea.L.hs_enc_encode_pair(ea.h, eb.h, xa.ctypes.data, xb.ctypes.data, ea.out.ctypes.data, eb.out.ctypes.data, ea.cap,
                                ea.nb.ctypes.data, eb.nb.ctypes.data, ret.ctypes.data)
    assert r == 0
    return tuple((bytes(e.out[:max(int(n), 0)]), (int(e.nb[0]), int(e.nb[1])), int(n)) for e, n in ((ea, ret[0]), (eb, ret[1])))


class SimDecoder:
    def __init__(self, use_md_index=0, framesize_ms=40, joint_hb=0):
        self.L = lib()
        self.h = self.L.hs_dec_create3(use_md_index, framesize_ms, joint_hb)
        self.pcm = np.zeros(16 * framesize_ms, np.int16)

    def decode(self, payload, nbytes, lostflag):
        buf = np.zeros(max(len(payload), 1), np.uint8)
        buf[:len(payload)] = np.frombuffer(payload, np.uint8)
        nb = np.array([nbytes[0], nbytes[1]], np.int16)
        r = self.L.hs_dec_decode(self.h, self.pcm.ctypes.data, buf.ctypes.data, len(buf), nb.ctypes.data, int(lostflag))
        return self.pcm.copy(), r

    def close(self):
        if self.h:
            self.L.hs_dec_destroy(self.h)
            self.h = None
