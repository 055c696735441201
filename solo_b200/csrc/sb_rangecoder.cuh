// solo_b200 -- SILK v1 range coder (16-bit range, 32-bit base, byte-wise renormalisation with backward
// carry propagation).  Reference: /root/reference/JC1_SDK_SRC_ARM/src/libSATECodec/SKP_Silk_range_coder.c.
#pragma once
#include "sb_common.cuh"
#include "sb_state.cuh"

namespace sb {

// SKP_Silk_range_enc_init (:249-259)
SB_HD void rc_enc_init(RangeEnc* rc, u8* buf, int bufLen) {
    rc->buf = buf; rc->bufLen = bufLen;
    rc->range_Q16 = 0x0000FFFF; rc->bufferIx = 0; rc->base_Q32 = 0; rc->error = 0;
}

// SKP_Silk_range_encoder (:31-98)
SB_FN void rc_encode(RangeEnc* rc, int data, const u16* prob) {
    if (rc->error) return;
    u32 base_Q32 = rc->base_Q32, range_Q16 = rc->range_Q16;
    i32 bufferIx = rc->bufferIx;
    u8* buffer = rc->buf;
    u32 low_Q16 = prob[data], high_Q16 = prob[data + 1];
    u32 base_tmp = base_Q32;
    base_Q32 += range_Q16 * low_Q16;
    u32 range_Q32 = range_Q16 * (high_Q16 - low_Q16);
    if (base_Q32 < base_tmp) {
        int ix = bufferIx;
        while ((++buffer[--ix]) == 0) {}
    }
    if (range_Q32 & 0xFF000000) {
        range_Q16 = range_Q32 >> 16;
    } else {
        if (range_Q32 & 0xFFFF0000) {
            range_Q16 = range_Q32 >> 8;
        } else {
            range_Q16 = range_Q32;
            if (bufferIx >= rc->bufLen) { rc->error = -1; return; }
            buffer[bufferIx++] = (u8)(base_Q32 >> 24);
            base_Q32 <<= 8;
        }
        if (bufferIx >= rc->bufLen) { rc->error = -1; return; }
        buffer[bufferIx++] = (u8)(base_Q32 >> 24);
        base_Q32 <<= 8;
    }
    rc->base_Q32 = base_Q32; rc->range_Q16 = range_Q16; rc->bufferIx = bufferIx;
}

// SKP_Silk_range_coder_get_length (:288-302)
SB_HD int rc_get_length(const RangeEnc* rc, int* nBytes) {
    int nBits = shl(rc->bufferIx, 3) + clz32((i32)(rc->range_Q16 - 1)) - 14;
    *nBytes = (nBits + 7) >> 3;
    return nBits;
}

// SKP_Silk_range_enc_wrap_up (:305-347)
SB_FN void rc_enc_wrap_up(RangeEnc* rc) {
    int nBytes;
    u32 base_Q24 = rc->base_Q32 >> 8;
    int bits_in_stream = rc_get_length(rc, &nBytes);
    int bits_to_store = bits_in_stream - shl(rc->bufferIx, 3);
    base_Q24 += 0x00800000u >> (bits_to_store - 1);
    base_Q24 &= 0xFFFFFFFFu << (24 - bits_to_store);
    if (base_Q24 & 0x01000000) {
        int ix = rc->bufferIx;
        while ((++(rc->buf[--ix])) == 0) {}
    }
    if (rc->bufferIx < rc->bufLen) {
        rc->buf[rc->bufferIx++] = (u8)(base_Q24 >> 16);
        if (bits_to_store > 8) {
            if (rc->bufferIx < rc->bufLen) rc->buf[rc->bufferIx++] = (u8)(base_Q24 >> 8);
        }
    }
    if (bits_in_stream & 7) {
        int mask = 0xFF >> (bits_in_stream & 7);
        if (nBytes - 1 < rc->bufLen) rc->buf[nBytes - 1] |= (u8)mask;
    }
}

// ---- decoder ---------------------------------------------------------------------------------------
struct RangeDec {
    u32 base_Q32, range_Q16;
    i32 bufferIx, error, bufLen;
    const u8* buf;  // payload; bytes at or beyond bufLen read as whatever the caller's padded copy holds
};

// SKP_Silk_range_dec_init (:262-285)
SB_HD void rc_dec_init(RangeDec* rc, const u8* buffer, int bufferLength) {
    rc->buf = buffer;
    if (bufferLength > MAX_PAYLOAD || bufferLength < 0) { rc->error = -8; return; }
    rc->bufLen = bufferLength;
    rc->bufferIx = 0;
    rc->base_Q32 = ((u32)buffer[0] << 24) | ((u32)buffer[1] << 16) | ((u32)buffer[2] << 8) | (u32)buffer[3];
    rc->range_Q16 = 0x0000FFFF;
    rc->error = 0;
}

// SKP_Silk_range_decoder (:115-231)
SB_FN void rc_decode(int* data, RangeDec* rc, const u16* prob, int probIx) {
    if (rc->error) { *data = 0; return; }
    u32 base_Q32 = rc->base_Q32, range_Q16 = rc->range_Q16;
    i32 bufferIx = rc->bufferIx;
    const u8* buffer = rc->buf + 4;
    u32 low_Q16, high_Q16 = prob[probIx];
    u32 base_tmp = range_Q16 * high_Q16;
    if (base_tmp > base_Q32) {
        while (1) {
            low_Q16 = prob[--probIx];
            base_tmp = range_Q16 * low_Q16;
            if (base_tmp <= base_Q32) break;
            high_Q16 = low_Q16;
            if (high_Q16 == 0) { rc->error = -2; *data = 0; return; }
        }
    } else {
        while (1) {
            low_Q16 = high_Q16;
            high_Q16 = prob[++probIx];
            base_tmp = range_Q16 * high_Q16;
            if (base_tmp > base_Q32) { probIx--; break; }
            if (high_Q16 == 0xFFFF) { rc->error = -2; *data = 0; return; }
        }
    }
    *data = probIx;
    base_Q32 -= range_Q16 * low_Q16;
    u32 range_Q32 = range_Q16 * (high_Q16 - low_Q16);
    if (range_Q32 & 0xFF000000) {
        range_Q16 = range_Q32 >> 16;
    } else {
        if (range_Q32 & 0xFFFF0000) {
            range_Q16 = range_Q32 >> 8;
            if (base_Q32 >> 24) { rc->error = -3; *data = 0; return; }
        } else {
            range_Q16 = range_Q32;
            if (((i32)base_Q32) >> 16) { rc->error = -3; *data = 0; return; }
            base_Q32 <<= 8;
            if (bufferIx < rc->bufLen) base_Q32 |= (u32)buffer[bufferIx++];
        }
        base_Q32 <<= 8;
        if (bufferIx < rc->bufLen) base_Q32 |= (u32)buffer[bufferIx++];
    }
    if (range_Q16 == 0) { rc->error = -4; *data = 0; return; }
    rc->base_Q32 = base_Q32; rc->range_Q16 = range_Q16; rc->bufferIx = bufferIx;
}

// SKP_Silk_range_coder_check_after_decoding (:350-371)
SB_FN void rc_check_after_decoding(RangeDec* rc) {
    int nBits = shl(rc->bufferIx, 3) + clz32((i32)(rc->range_Q16 - 1)) - 14;
    int nBytes = (nBits + 7) >> 3;
    if (nBytes - 1 >= rc->bufLen) { rc->error = -5; return; }
    if (nBits & 7) {
        int mask = 0xFF >> (nBits & 7);
        if ((rc->buf[nBytes - 1] & mask) != mask) { rc->error = -5; return; }
    }
}

}  // namespace sb
