// Host build of the per-stream kernel source (solo_b200/csrc/*.cuh compiled by g++).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only container check the kernel logic bit-for-bit against the
// compiled reference (oracle/_ref).  Never linked into libsolo_b200.so.
#include "../../solo_b200/csrc/sb_enc.cuh"
#include <stdlib.h>

extern "C" {
struct HsEnc { sb::EncState st; sb::EncPacketWork w; };
void* hs_enc_create(int rate, int dtx, int mdi) {
    HsEnc* h = (HsEnc*)calloc(1, sizeof(HsEnc));
    sb::enc_state_init(&h->st, rate, dtx, mdi);
    return h;
}
int hs_enc_encode(void* p, const short* pcm, unsigned char* out, int cap, short* nb) {
    HsEnc* h = (HsEnc*)p;
    return sb::enc_packet(&h->st, &h->w, pcm, out, cap, nb);
}
void hs_enc_destroy(void* p) { free(p); }
int hs_enc_state_size() { return (int)sizeof(sb::EncState); }
int hs_enc_work_size() { return (int)sizeof(sb::EncPacketWork); }
void* hs_enc_state(void* p) { return &((HsEnc*)p)->st; }
void* hs_enc_ctrl(void* p) { return &((HsEnc*)p)->w.f.c; }
}
