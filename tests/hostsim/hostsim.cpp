// Host build of the per-stream kernel source (solo_b200/csrc/*.cuh compiled by g++).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only container check the kernel logic bit-for-bit against the
// compiled reference (oracle/_ref).  Never linked into libsolo_b200.so.
// With -DSB_EMU the analysis stage runs in its cooperative form on 32 OS threads per stream (sb_par.cuh), so that
// races / missing barriers of the warp-per-stream kernel can be caught without a GPU.
#ifdef SB_EMU
// 32 lanes of one stream as 32 fibers on one OS thread, run round-robin between barriers: lane 0 runs to its next barrier,
// then lane 1, ... so a lane that reads what another lane has not written yet (missing SB_SYNC) or that takes a different
// number of barriers (divergent collective) shows up deterministically.
#include <ucontext.h>
#include <stdio.h>
#include <stdlib.h>
#include <functional>
#include "../../solo_b200/csrc/sb_common.cuh"
namespace sb { namespace emu {
int lane = 0;
static ucontext_t ctx[32], main_ctx;
static char* stacks[32];
static long long scr[32];
static long nbar[32];
static std::function<void()> body;
void barrier() {
    const int me = lane;
    nbar[me]++;
    lane = (me + 1) & 31;
    swapcontext(&ctx[me], &ctx[lane]);
    lane = me;
    if (me == 0)      // a full round has passed: every lane must have arrived at its barrier number nbar[0]
        for (int i = 1; i < 32; i++)
            if (nbar[i] != nbar[0]) { fprintf(stderr, "emu: divergent collective -- lane %d is at barrier %ld, lane 0 at %ld\n", i, nbar[i], nbar[0]); abort(); }
}
long long* scratch() { return scr; }
static void entry() { body(); }
// run f on 32 lanes (f reads sb::emu::lane)
static void run32(const std::function<void()>& f) {
    body = f;
    for (int i = 0; i < 32; i++) {
        if (!stacks[i]) stacks[i] = (char*)malloc(1 << 20);
        getcontext(&ctx[i]);
        ctx[i].uc_stack.ss_sp = stacks[i];
        ctx[i].uc_stack.ss_size = 1 << 20;
        ctx[i].uc_link = i < 31 ? &ctx[i + 1] : &main_ctx;
        makecontext(&ctx[i], (void (*)())entry, 0);
        nbar[i] = 0;
    }
    // a lane that returns falls through (uc_link) into the next lane's saved context; `lane` is restored inside barrier()
    lane = 0;
    swapcontext(&main_ctx, &ctx[0]);
    for (int i = 1; i < 32; i++)
        if (nbar[i] != nbar[0]) { fprintf(stderr, "emu: lane %d took %ld barriers, lane 0 took %ld (divergent collective)\n", i, nbar[i], nbar[0]); abort(); }
    lane = 0;
}

// ---- a second scheduler for code whose collectives name a lane mask (the quantiser kernel: two 16-lane groups that may take
//      different paths between the full-warp sections).  barrier_mask(m) really waits: a lane spins (yielding round-robin to
//      the other lanes) until every lane of m has arrived at a barrier with the same mask.  Lanes finish at different times. ----
static bool done_[32];
static struct MaskBar { unsigned mask, arrived; unsigned long phase; } mbar[8];
static int n_mbar;
static long spins;
static void yield_next() {
    const int me = lane;
    for (int k = 1; k <= 32; k++) {
        const int nx = (me + k) & 31;
        if (!done_[nx]) { if (nx == me) return; lane = nx; swapcontext(&ctx[me], &ctx[nx]); lane = me; return; }
    }
}
void barrier_mask(unsigned m) {
    const int me = lane;
    MaskBar* e = nullptr;
    for (int i = 0; i < n_mbar; i++) if (mbar[i].mask == m) e = &mbar[i];
    if (!e) { if (n_mbar == 8) { fprintf(stderr, "emu: too many distinct lane masks\n"); abort(); } e = &mbar[n_mbar++]; e->mask = m; e->arrived = 0; e->phase = 0; }
    if (!((m >> me) & 1)) { fprintf(stderr, "emu: lane %d at a barrier whose mask %08x does not name it\n", me, m); abort(); }
    const unsigned long my_phase = e->phase;
    e->arrived |= 1u << me;
    if (e->arrived == m) { e->arrived = 0; e->phase++; spins = 0; }
    while (e->phase == my_phase) {
        if (++spins > 100000000L) { fprintf(stderr, "emu: deadlock -- lanes %08x of mask %08x never arrived (divergent collective)\n", m & ~e->arrived, m); abort(); }
        yield_next();
    }
}
static void entry_m() {
    body();
    const int me = lane;
    done_[me] = true;
    for (int k = 1; k < 32; k++) { const int nx = (me + k) & 31; if (!done_[nx]) { lane = nx; setcontext(&ctx[nx]); } }
    setcontext(&main_ctx);
}
static void run32m(const std::function<void()>& f) {
    body = f;
    n_mbar = 0; spins = 0;
    for (int i = 0; i < 32; i++) {
        if (!stacks[i]) stacks[i] = (char*)malloc(1 << 20);
        getcontext(&ctx[i]);
        ctx[i].uc_stack.ss_sp = stacks[i];
        ctx[i].uc_stack.ss_size = 1 << 20;
        ctx[i].uc_link = &main_ctx;
        makecontext(&ctx[i], (void (*)())entry_m, 0);
        done_[i] = false;
    }
    lane = 0;
    swapcontext(&main_ctx, &ctx[0]);
    lane = 0;
}
} }
#endif
#include "../../solo_b200/csrc/sb_enc.cuh"
#include "../../solo_b200/csrc/sb_dec.cuh"
#ifdef SB_EMU
#include "../../solo_b200/csrc/sb_coop.cuh"
// ---- the quantiser kernel's device code (sb_nsq_warp.cuh) under the same 32-fibre emulation: one stream per "warp"
//      (SB_NSQ_GW = 32), the CUDA warp intrinsics it uses as shims over the emulation's exchange buffer ----
namespace {
struct EmuThreadIdx { struct X { operator int() const { return sb::emu::lane; } } x; } threadIdx;
static long long xbuf[32];     // exchange line of the shims (one slot per lane)
template <class T> inline T emu_xchg(unsigned m, T v, int src) {
    xbuf[sb::emu::lane] = (long long)v;
    sb::emu::barrier_mask(m);
    const T r = (T)xbuf[src & 31];
    sb::emu::barrier_mask(m);
    return r;
}
template <class T> inline T __shfl_sync(unsigned m, T v, int src, int w = 32) { return emu_xchg<T>(m, v, (sb::emu::lane & ~(w - 1)) | (src & (w - 1))); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, int d, int w = 32) {
    const int l = sb::emu::lane;
    return emu_xchg<T>(m, v, ((l & (w - 1)) + d < w) ? l + d : l);
}
inline unsigned __ballot_sync(unsigned m, bool p) {
    xbuf[sb::emu::lane] = p ? 1 : 0;
    sb::emu::barrier_mask(m);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((m >> i) & 1) r |= (unsigned)xbuf[i] << i;
    sb::emu::barrier_mask(m);
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline int __reduce_max_sync(unsigned m, int v) {
    xbuf[sb::emu::lane] = v;
    sb::emu::barrier_mask(m);
    int r = v;
    for (int i = 0; i < 32; i++) if (((m >> i) & 1) && (int)xbuf[i] > r) r = (int)xbuf[i];
    sb::emu::barrier_mask(m);
    return r;
}
inline void __syncwarp(unsigned m = 0xffffffffu) { sb::emu::barrier_mask(m); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __mulhi(int a, int b) { return (int)(((long long)a * (long long)b) >> 32); }
}
#ifdef SB_EMU_GW16
#define SB_NSQ_GW 16     // the kernel's packing: two lane groups per warp; here the second group shadows the first one's stream
#else                    // (what the kernel does with the last stream of an odd batch), so both take the same path
#define SB_NSQ_GW 32
#endif
#include "../../solo_b200/csrc/sb_nsq_warp.cuh"
static int g_emu_nsq = 0;
extern "C" void hs_set_emu_nsq(int on) { g_emu_nsq = on; }
#endif
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
#ifdef SB_EMU
// the device pipeline up to the quantiser: band split (kernel A0), VAD, core analysis (cooperative), the thread-per-stream
// kernels behind it, high-band analysis (cooperative)
static void emu_front(HsEnc* h, const short* pcm) {
    static sb::CoopWork cw;
    const int nf = h->st.frames_per_packet;
    sb::qmf_decomp(pcm, h->w.a.low, h->w.a.high, h->st.qmf_mem, nf * 2 * sb::FRAME);
    for (int i = 0; i < nf * sb::FRAME; i++) cw.low[i] = h->w.a.low[i];
    sb::vad_packet(&h->st.vad, h->w.a.low, nf, h->w.scr.vad_sa_Q8, h->w.scr.vad_quality_Q15, h->w.scr.vad_tilt_Q15);   // the VAD kernel
    sb::emu::run32([=]() { sb::c_enc_packet_analysis(&h->st, &cw, &h->w.scr); });
    for (int f = 0; f < nf; f++) for (int k = 0; k < sb::NB_SUBFR; k++) sb::shape_post_window(&h->w.scr, f, k);   // shaping-filter kernel
    sb::gains_packet(&h->st, &h->w.scr, nf);                                                                      // gain kernel
    sb::prefilter_packet(&h->st, &h->w.scr, nf);                                                                  // prefilter kernel
    static sb::HbScr hs;
    sb::emu::run32([=]() {
        for (int f = 0; f < nf * sb::FRAME / h->st.hb_frame; f++) {
            if (h->st.hb_frame == sb::HB_FRAME) sb::c_hb_analyse_frame<sb::HB_FRAME>(&h->st, &hs, h->w.a.high + f * h->st.hb_frame, &h->w.scr.hb_lsp_idx[f], h->w.scr.hb_nrg0[f]);
            else sb::c_hb_analyse_frame<2 * sb::HB_FRAME>(&h->st, &hs, h->w.a.high + f * h->st.hb_frame, &h->w.scr.hb_lsp_idx[f], h->w.scr.hb_nrg0[f]);
        }
    });
}
// the quantiser kernel's code for one warp: lane group g works on stream hh[g] (one group with SB_NSQ_GW = 32)
static void emu_nsq(HsEnc* const* hh) {
    static sb::NsqSmem S[2];
    const int nf = hh[0]->st.frames_per_packet;
    for (int f = 0; f < nf; f++)
        sb::emu::run32m([=]() {
            const int g = sb::emu::lane / SB_NSQ_GW;
            HsEnc* h = hh[g];
            sb::EncScratch* scr = &h->w.scr;
            sb::nsq_del_dec_warp(S[g], h->st.nsq, &scr->c[f], scr->xfw[f], scr->q_md[f][0], scr->q_md[f][1], scr->r16[f], &scr->nsq_rand[0][0][0]);
        });
}
#endif
int hs_enc_encode(void* p, const short* pcm, unsigned char* out, int cap, short* nb) {
    HsEnc* h = (HsEnc*)p;
#ifdef SB_EMU
    emu_front(h, pcm);
    if (g_emu_nsq) {      // second lane group (SB_NSQ_GW = 16): shadows the same stream, as the kernel does for an odd batch
        HsEnc* hh[2] = {h, h};
        emu_nsq(hh);
        return sb::enc_packet_finish(&h->st, &h->w.scr, h->w.rcbuf, out, cap, nb);
    }
    return sb::enc_packet_quantise_and_code(&h->st, &h->w, out, cap, nb);
#else
    return sb::enc_packet(&h->st, &h->w, pcm, out, cap, nb);
#endif
}
// two DIFFERENT streams through one emulated quantiser warp (SB_NSQ_GW = 16 build): the lane groups take their own paths
// between the full-warp sample loops.  Both encoders must have the same packet length.  ret[g] = bytes of stream g.
int hs_enc_encode_pair(void* pa, void* pb, const short* pcm_a, const short* pcm_b, unsigned char* out_a, unsigned char* out_b, int cap,
                       short* nb_a, short* nb_b, int* ret) {
#if defined(SB_EMU) && defined(SB_EMU_GW16)
    HsEnc* hh[2] = {(HsEnc*)pa, (HsEnc*)pb};
    if (hh[0]->st.frames_per_packet != hh[1]->st.frames_per_packet) return -1;
    emu_front(hh[0], pcm_a);
    emu_front(hh[1], pcm_b);
    emu_nsq(hh);
    ret[0] = sb::enc_packet_finish(&hh[0]->st, &hh[0]->w.scr, hh[0]->w.rcbuf, out_a, cap, nb_a);
    ret[1] = sb::enc_packet_finish(&hh[1]->st, &hh[1]->w.scr, hh[1]->w.rcbuf, out_b, cap, nb_b);
    return 0;
#else
    return -1;
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
