// Host build of the per-stream kernel source (solo_b200/csrc/*.cuh compiled by g++).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only container check the kernel logic bit-for-bit against the
// compiled reference (oracle/_ref).  Never linked into libsolo_b200.so.
// With -DSB_EMU the analysis stage runs in its cooperative form on 32 OS threads per stream (sb_par.cuh), so that
// races / missing barriers of the warp-per-stream kernel can be caught without a GPU.
#ifdef SB_EMU
#include <pthread.h>
#include <thread>
#include <vector>
#include "../../solo_b200/csrc/sb_common.cuh"
namespace sb { namespace emu {
thread_local int lane = 0;
static pthread_barrier_t bar;
static long long scr[32];
static bool inited = false;
void barrier() { pthread_barrier_wait(&bar); }
long long* scratch() { return scr; }
static void init() { if (!inited) { pthread_barrier_init(&bar, nullptr, 32); inited = true; } }
} }
#endif
#include "../../solo_b200/csrc/sb_enc.cuh"
#include "../../solo_b200/csrc/sb_dec.cuh"
#include <stdlib.h>

extern "C" {
struct HsEnc { sb::EncState st; sb::EncPacketWork w; };
void* hs_enc_create3(int rate, int dtx, int mdi, int framesize_ms, int joint_hb) {
    HsEnc* h = (HsEnc*)calloc(1, sizeof(HsEnc));
    sb::enc_state_init(&h->st, rate, dtx, mdi, framesize_ms, joint_hb);
    h->w.a.nlsf_fast = nullptr;
    return h;
}
void* hs_enc_create2(int rate, int dtx, int mdi, int framesize_ms) { return hs_enc_create3(rate, dtx, mdi, framesize_ms, 0); }
void* hs_enc_create(int rate, int dtx, int mdi) { return hs_enc_create2(rate, dtx, mdi, 40); }
int hs_enc_encode(void* p, const short* pcm, unsigned char* out, int cap, short* nb) {
    HsEnc* h = (HsEnc*)p;
#ifdef SB_EMU
    sb::emu::init();
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++)
        th.emplace_back([=]() { sb::emu::lane = l; sb::enc_packet_analysis(&h->st, &h->w.a, pcm, &h->w.scr); });
    for (auto& t : th) t.join();
    sb::emu::lane = 0;
    return sb::enc_packet_quantise_and_code(&h->st, &h->w, out, cap, nb);
#else
    return sb::enc_packet(&h->st, &h->w, pcm, out, cap, nb);
#endif
}
void hs_enc_destroy(void* p) { free(p); }
int hs_is_emu() {
#ifdef SB_EMU
    return 1;
#else
    return 0;
#endif
}
int hs_enc_state_size() { return (int)sizeof(sb::EncState); }
int hs_enc_work_size() { return (int)sizeof(sb::EncPacketWork); }
void* hs_enc_state(void* p) { return &((HsEnc*)p)->st; }
void* hs_enc_ctrl(void* p) { return &((HsEnc*)p)->w.scr.c[1]; }

struct HsDec { sb::DecState st; sb::DecPacketWork w; sb::DecStale stale; };
void* hs_dec_create3(int mdi, int framesize_ms, int joint_hb) {
    HsDec* h = (HsDec*)calloc(1, sizeof(HsDec));
    sb::dec_state_init(&h->st, mdi, framesize_ms, joint_hb);
    return h;
}
void* hs_dec_create2(int mdi, int framesize_ms) { return hs_dec_create3(mdi, framesize_ms, 0); }
void* hs_dec_create(int mdi) { return hs_dec_create2(mdi, 40); }
// same calling convention as AGR_Sate_Decoder_Decode (payload pre-trimmed by the caller); nb is not modified
int hs_dec_decode(void* p, short* pcm, const unsigned char* bits, int cap, const short* nb, int lostflag) {
    HsDec* h = (HsDec*)p;
    return sb::dec_packet(&h->st, &h->w, pcm, bits, cap, nb, lostflag, &h->stale);
}
void hs_dec_destroy(void* p) { free(p); }
int hs_dec_state_size() { return (int)sizeof(sb::DecState); }
int hs_dec_work_size() { return (int)sizeof(sb::DecPacketWork); }
}
