"""ctypes binding of libsolo_b200.so.

Mirrors the reference interface for the hot path:
  * ``SoloEncoder`` / ``SoloDecoder``  -- the six ``AGR_Sate_*`` entry points (one stream per handle), same argument
    meaning and return conventions as /root/reference/JC1_SDK_SRC_ARM/interface/AGR_JC1_SDK_API.h:33-64;
  * ``EncoderBatch`` / ``DecoderBatch`` -- the batched extension (``include/solo_b200.h``), N streams per call.
Buffers are numpy arrays (host entry points) or raw device pointers (ints) for the ``*_device`` entry points.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SOLO_B200_LIB: load another build of the same CUDA library (kernel-variant experiments); never a CPU path
LIB_PATH = os.environ.get("SOLO_B200_LIB") or os.path.join(_HERE, "libsolo_b200.so")
PACKET = 640


class SoloError(RuntimeError):
    pass


class EncCtrl(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mode", "targetRate_bps", "samplerate", "dtx_enable", "framesize_ms",
                                         "joint_enable", "joint_mode", "useMDIndex")]


class DecCtrl(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("packetLoss_perc", "samplerate", "framesize_ms", "joint_enable",
                                         "joint_mode", "useMDIndex")]


_lib = None


def lib():
    """Load libsolo_b200.so (fails loudly when it has not been built: there is no CPU implementation)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SoloError("%s missing: run `python -m solo_b200.build` (needs nvcc); no CPU fallback exists" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32p = C.c_void_p, C.c_void_p
        L.solo_b200_last_error.restype = C.c_char_p
        L.solo_b200_set_chunks.argtypes = [C.c_int]
        L.solo_b200_kernel_launches.restype = C.c_longlong
        L.solo_b200_enc_batch_create.restype = vp
        L.solo_b200_enc_batch_create.argtypes = [C.c_int, C.POINTER(EncCtrl), C.c_int]
        L.solo_b200_enc_batch_encode_host.argtypes = [vp, vp, vp, C.c_int, vp]
        L.solo_b200_enc_batch_encode_device.argtypes = [vp, vp, vp, C.c_int, vp, vp]
        L.solo_b200_enc_batch_destroy.argtypes = [vp]
        L.solo_b200_dec_batch_create.restype = vp
        L.solo_b200_dec_batch_create.argtypes = [C.c_int, C.POINTER(DecCtrl), C.c_int]
        L.solo_b200_dec_batch_decode_host.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
        L.solo_b200_dec_batch_decode_device.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.solo_b200_dec_batch_destroy.argtypes = [vp]
        for f in ("solo_b200_enc_batch_export_state", "solo_b200_enc_batch_import_state", "solo_b200_dec_batch_export_state",
                  "solo_b200_dec_batch_import_state"):
            getattr(L, f).argtypes = [vp, C.c_int, vp]
        pp, ip = C.POINTER(vp), C.POINTER(C.c_int)
        L.solo_b200_split_packet.argtypes = [vp, vp, pp, ip, pp, ip]
        L.solo_b200_merge_packets.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp]
        L.solo_b200_bitfile_pack.argtypes = [vp, vp, vp, C.c_int]
        L.solo_b200_bitfile_unpack.argtypes = [vp, C.c_int, pp, vp]
        L.solo_b200_apply_loss_device.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        L.solo_b200_enable_peer_access.argtypes = [C.c_int, C.c_int]
        L.solo_b200_enc_state_bytes.restype = C.c_int
        L.solo_b200_dec_state_bytes.restype = C.c_int
        L.solo_b200_profile_enable.argtypes = [C.c_int]
        L.solo_b200_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.AGR_Sate_Encoder_Init.restype = vp
        L.AGR_Sate_Encoder_Init.argtypes = [C.POINTER(EncCtrl)]
        L.AGR_Sate_Encoder_Encode.restype = C.c_int32
        L.AGR_Sate_Encoder_Encode.argtypes = [vp, vp, vp, C.c_int32, vp]
        L.AGR_Sate_Encoder_Uninit.argtypes = [vp]
        L.AGR_Sate_Decoder_Init.restype = vp
        L.AGR_Sate_Decoder_Init.argtypes = [C.POINTER(DecCtrl)]
        L.AGR_Sate_Decoder_Decode.restype = C.c_int32
        L.AGR_Sate_Decoder_Decode.argtypes = [vp, vp, vp, vp, vp, C.c_int32]
        L.AGR_Sate_Decoder_Uninit.argtypes = [vp]
        _lib = L
    return _lib


def _err():
    return lib().solo_b200_last_error().decode("utf-8", "replace")


def kernel_launches():
    return int(lib().solo_b200_kernel_launches())


def state_bytes():
    return int(lib().solo_b200_enc_state_bytes()), int(lib().solo_b200_dec_state_bytes())


def profile_enable(on=True):
    lib().solo_b200_profile_enable(1 if on else 0)


def profile_read():
    """{kernel: (total_ms, launches)} since the last read."""
    t = (C.c_double * 4)()
    n = (C.c_longlong * 4)()
    lib().solo_b200_profile_read(t, n)
    names = ("enc_analysis", "enc_nsq", "enc_finish", "decode")
    return {k: (t[i], n[i]) for i, k in enumerate(names)}


class SoloEncoder:
    """One stream through AGR_Sate_Encoder_Init / Encode / Uninit."""

    def __init__(self, rate=13600, dtx=0, use_md_index=0, samplerate=16000, framesize_ms=40, joint_enable=0, joint_mode=0):
        self.ctrl = EncCtrl(2, rate, samplerate, dtx, framesize_ms, joint_enable, joint_mode, use_md_index)
        self.h = lib().AGR_Sate_Encoder_Init(C.byref(self.ctrl))
        if not self.h:
            raise SoloError("AGR_Sate_Encoder_Init returned NULL: " + _err())
        self._bits = np.zeros(1024, np.uint8)
        self._nb = np.zeros(6, np.int16)
        self.samples = 16 * framesize_ms

    def encode(self, pcm640, bufsize=1024):
        pcm = np.ascontiguousarray(pcm640, dtype=np.int16)
        assert pcm.size == self.samples
        self._nb[:] = 0
        n = lib().AGR_Sate_Encoder_Encode(self.h, pcm.ctypes.data, self._bits.ctypes.data, bufsize, self._nb.ctypes.data)
        return bytes(self._bits[:max(n, 0)]), (int(self._nb[0]), int(self._nb[1])), n

    def close(self):
        if getattr(self, "h", None):
            lib().AGR_Sate_Encoder_Uninit(self.h)
            self.h = None

    __del__ = close


class SoloDecoder:
    """One stream through AGR_Sate_Decoder_Init / Decode / Uninit (payload pre-trimmed by the caller)."""

    def __init__(self, use_md_index=0, samplerate=16000, framesize_ms=40, joint_enable=0, joint_mode=0):
        self.ctrl = DecCtrl(0, samplerate, framesize_ms, joint_enable, joint_mode, use_md_index)
        self.h = lib().AGR_Sate_Decoder_Init(C.byref(self.ctrl))
        if not self.h:
            raise SoloError("AGR_Sate_Decoder_Init returned NULL: " + _err())
        self._out = np.zeros(960, np.int16)
        self._ns = C.c_int16(0)
        self.samples = 16 * framesize_ms

    def decode(self, payload, nbytes, lostflag):
        buf = np.zeros(1040, np.uint8)
        buf[:len(payload)] = np.frombuffer(bytes(payload), np.uint8)
        nb = np.array([nbytes[0], nbytes[1], 0, 0, 0, 0], np.int16)
        ret = lib().AGR_Sate_Decoder_Decode(self.h, self._out.ctypes.data, C.byref(self._ns), buf.ctypes.data, nb.ctypes.data, int(lostflag))
        self.last_nbytes = (int(nb[0]), int(nb[1]))
        self.last_nsamples = int(self._ns.value)
        return self._out[:self.samples].copy(), ret

    def close(self):
        if getattr(self, "h", None):
            lib().AGR_Sate_Decoder_Uninit(self.h)
            self.h = None

    __del__ = close


class EncoderBatch:
    """N encoder streams resident on one GPU; one call = one 40 ms packet for every stream."""

    def __init__(self, n, rate=13600, dtx=0, use_md_index=0, device=0, framesize_ms=40, joint_hb=0):
        """joint_hb=1: the reference's joint mode 1 (one 40 ms high-band frame per packet)."""
        self.n = int(n)
        self.samples = 16 * framesize_ms          # samples per packet and stream (row length of the PCM matrix)
        self.ctrl = EncCtrl(2, rate, 16000, dtx, framesize_ms, 1 if joint_hb else 0, 1 if joint_hb else 0, use_md_index)
        self.h = lib().solo_b200_enc_batch_create(self.n, C.byref(self.ctrl), device)
        if not self.h:
            raise SoloError("solo_b200_enc_batch_create failed: " + _err())

    def encode(self, pcm, cap=256, bits=None, nbytes=None):
        """pcm: int16 [N, 640] host array -> (bits uint8 [N, cap], nbytes int16 [N, 2])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        assert pcm.shape == (self.n, self.samples)
        if bits is None:
            bits = np.empty((self.n, cap), np.uint8)
        if nbytes is None:
            nbytes = np.empty((self.n, 2), np.int16)
        r = lib().solo_b200_enc_batch_encode_host(self.h, pcm.ctypes.data, bits.ctypes.data, cap, nbytes.ctypes.data)
        if r:
            raise SoloError("encode_host failed (%d): %s" % (r, _err()))
        return bits, nbytes

    def encode_ptr(self, pcm_ptr, bits_ptr, cap, nbytes_ptr):
        """Host entry point on raw addresses (e.g. pinned torch tensors' data_ptr())."""
        r = lib().solo_b200_enc_batch_encode_host(self.h, pcm_ptr, bits_ptr, cap, nbytes_ptr)
        if r:
            raise SoloError("encode_host failed (%d): %s" % (r, _err()))

    def export_state(self, idx):
        """Opaque blob with the complete state of stream idx (checkpoint / migration)."""
        blob = np.zeros(lib().solo_b200_enc_state_bytes(), np.uint8)
        if lib().solo_b200_enc_batch_export_state(self.h, int(idx), blob.ctypes.data):
            raise SoloError(_err())
        return blob

    def import_state(self, idx, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        if blob.size != lib().solo_b200_enc_state_bytes() or lib().solo_b200_enc_batch_import_state(self.h, int(idx), blob.ctypes.data):
            raise SoloError("import_state: " + _err())

    def encode_device(self, d_pcm, d_bits, cap, d_nbytes, stream=0):
        r = lib().solo_b200_enc_batch_encode_device(self.h, d_pcm, d_bits, cap, d_nbytes, stream)
        if r:
            raise SoloError("encode_device failed (%d): %s" % (r, _err()))

    def close(self):
        if getattr(self, "h", None):
            lib().solo_b200_enc_batch_destroy(self.h)
            self.h = None

    __del__ = close


class DecoderBatch:
    """N decoder streams resident on one GPU."""

    def __init__(self, n, use_md_index=0, device=0, framesize_ms=40, joint_hb=0):
        self.n = int(n)
        self.samples = 16 * framesize_ms
        self.ctrl = DecCtrl(0, 16000, framesize_ms, 1 if joint_hb else 0, 1 if joint_hb else 0, use_md_index)
        self.h = lib().solo_b200_dec_batch_create(self.n, C.byref(self.ctrl), device)
        if not self.h:
            raise SoloError("solo_b200_dec_batch_create failed: " + _err())

    def decode(self, bits, nbytes, lostflag, pcm=None, ret=None):
        """bits uint8 [N, cap], nbytes int16 [N, 2], lostflag int32 [N] -> (pcm int16 [N, 640], ret int32 [N])."""
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        nbytes = np.ascontiguousarray(nbytes, dtype=np.int16)
        lostflag = np.ascontiguousarray(lostflag, dtype=np.int32)
        assert bits.shape[0] == self.n and nbytes.shape == (self.n, 2) and lostflag.shape == (self.n,)
        if pcm is None:
            pcm = np.zeros((self.n, self.samples), np.int16)
        if ret is None:
            ret = np.zeros(self.n, np.int32)
        r = lib().solo_b200_dec_batch_decode_host(self.h, pcm.ctypes.data, bits.ctypes.data, bits.shape[1], nbytes.ctypes.data,
                                                  lostflag.ctypes.data, ret.ctypes.data)
        if r:
            raise SoloError("decode_host failed (%d): %s" % (r, _err()))
        return pcm, ret

    def decode_ptr(self, pcm_ptr, bits_ptr, cap, nbytes_ptr, flags_ptr, ret_ptr):
        r = lib().solo_b200_dec_batch_decode_host(self.h, pcm_ptr, bits_ptr, cap, nbytes_ptr, flags_ptr, ret_ptr)
        if r:
            raise SoloError("decode_host failed (%d): %s" % (r, _err()))

    def export_state(self, idx):
        blob = np.zeros(lib().solo_b200_dec_state_bytes(), np.uint8)
        if lib().solo_b200_dec_batch_export_state(self.h, int(idx), blob.ctypes.data):
            raise SoloError(_err())
        return blob

    def import_state(self, idx, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        if blob.size != lib().solo_b200_dec_state_bytes() or lib().solo_b200_dec_batch_import_state(self.h, int(idx), blob.ctypes.data):
            raise SoloError("import_state: " + _err())

    def decode_device(self, d_pcm, d_bits, cap, d_nbytes, d_flags, d_ret=0, stream=0):
        r = lib().solo_b200_dec_batch_decode_device(self.h, d_pcm, d_bits, cap, d_nbytes, d_flags, d_ret, stream)
        if r:
            raise SoloError("decode_device failed (%d): %s" % (r, _err()))

    def close(self):
        if getattr(self, "h", None):
            lib().solo_b200_dec_batch_destroy(self.h)
            self.h = None

    __del__ = close


def set_chunks(n):
    """Number of stream groups a packet wave is pipelined as (see include/solo_b200.h); 1 = plain single launches."""
    lib().solo_b200_set_chunks(int(n))


# ---- packet framing (host functions of the library; no GPU involved) --------------------------------------------------
def split_packet(bits, nbytes):
    """(description 1, description 2) = (MD1, MD2 + 8 high-band bytes) of one payload row; (b"", b"") for a DTX row."""
    buf = np.ascontiguousarray(np.frombuffer(bytes(bits), np.uint8))
    nb = np.array([nbytes[0], nbytes[1]], np.int16)
    p1, p2, n1, n2 = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
    if lib().solo_b200_split_packet(buf.ctypes.data, nb.ctypes.data, C.byref(p1), C.byref(n1), C.byref(p2), C.byref(n2)):
        raise SoloError("split_packet: inconsistent length fields")
    o1 = (p1.value or buf.ctypes.data) - buf.ctypes.data
    o2 = (p2.value or buf.ctypes.data) - buf.ctypes.data
    return bytes(buf[o1:o1 + n1.value]), bytes(buf[o2:o2 + n2.value])


def merge_packets(p1, p2, cap=1024):
    """(payload row, (n0, n1), lostflag) for the decoder from whichever descriptions arrived (None / b"" = lost)."""
    a = np.frombuffer(p1, np.uint8).copy() if p1 else None
    b = np.frombuffer(p2, np.uint8).copy() if p2 else None
    out = np.zeros(cap, np.uint8)
    nb = np.zeros(2, np.int16)
    flag = C.c_int32()
    r = lib().solo_b200_merge_packets(a.ctypes.data if a is not None else None, len(a) if a is not None else 0,
                                      b.ctypes.data if b is not None else None, len(b) if b is not None else 0,
                                      out.ctypes.data, cap, nb.ctypes.data, C.byref(flag))
    if r:
        raise SoloError("merge_packets: packets do not fit")
    return bytes(out[:max(int(nb[0]), 0)]), (int(nb[0]), int(nb[1])), int(flag.value)


def bitfile_pack(bits, nbytes):
    """One record of the reference CLI's .bit file."""
    buf = np.frombuffer(bytes(bits) + b"\0", np.uint8).copy()
    nb = np.array([nbytes[0], nbytes[1]], np.int16)
    out = np.zeros(4 + max(int(nbytes[0]), 0), np.uint8)
    n = lib().solo_b200_bitfile_pack(buf.ctypes.data, nb.ctypes.data, out.ctypes.data, out.size)
    if n < 0:
        raise SoloError("bitfile_pack")
    return bytes(out[:n])


def bitfile_unpack(data, offset=0):
    """(payload, (n0, n1), next offset) of the record starting at `offset`."""
    buf = np.frombuffer(data, np.uint8)
    view = np.ascontiguousarray(buf[offset:])
    nb = np.zeros(2, np.int16)
    p = C.c_void_p()
    n = lib().solo_b200_bitfile_unpack(view.ctypes.data, view.size, C.byref(p), nb.ctypes.data)
    if n < 0:
        raise SoloError("bitfile_unpack: truncated record")
    return bytes(view[4:n]), (int(nb[0]), int(nb[1])), offset + n


def enable_peer_access(device, peer_device):
    if lib().solo_b200_enable_peer_access(int(device), int(peer_device)):
        raise SoloError(_err())


def apply_loss_device(d_bits_in, d_nbytes_in, d_lostflag, d_bits_out, d_nbytes_out, cap, n, stream=0):
    if lib().solo_b200_apply_loss_device(d_bits_in, d_nbytes_in, d_lostflag, d_bits_out, d_nbytes_out, cap, n, stream):
        raise SoloError(_err())
