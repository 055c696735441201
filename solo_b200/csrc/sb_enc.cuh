// solo_b200 -- one SOLO packet (40 ms, 16 kHz) through the encoder: QMF split, two SILK frames with the
// MD noise-shaping quantiser, two high-band frames, payload assembly [MD1 | MD2 | HB].
// Reference: /root/reference/JC1_SDK_SRC_ARM/src/libBWE/AGR_BWE_encode_frame_FIX.c:8-174, AGR_BWE_qmf.c:38-80,
// AGR_BWE_find_HB_LPC_FIX.c:4-49, AGR_BWE_quant_highband.c:23-147, AGR_BWE_SDK_API.c:11-152,
// libSATECodec/SKP_Silk_enc_API.c:127-273, SKP_Silk_encode_frame_FIX.c:34-327, SKP_Silk_control_codec_FIX.c:232-389.
#pragma once
#include "sb_enc_entropy.cuh"
#include "sb_enc_front.cuh"
#include "sb_enc_pred.cuh"
#include "sb_enc_shape.cuh"
#include "sb_nsq.cuh"

namespace sb {

// ---- Init: AGR_Sate_Encoder_Init + SKP_Silk_init_encoder_FIX + first SKP_Silk_control_encoder_FIX --------------
// Returns 0, or -1 for a configuration the reference rejects / this build does not cover.
SB_FN int enc_state_init(EncState* st, i32 targetRate_bps, i32 dtx_enable, i32 useMDIndex, i32 framesize_ms = 40, i32 joint_hb = 0) {
    memset(st, 0, sizeof(EncState));
    st->frames_per_packet = framesize_ms == 20 ? 1 : 2;
    st->hb_frame = joint_hb ? 320 : HB_FRAME;
    if (targetRate_bps <= 0) targetRate_bps = 15600;
    i32 silk_rate = targetRate_bps - (joint_hb ? 800 : 1600);  // (1600 * 20) / bwe_framesize_ms (AGR_BWE_SDK_API.c:118)
    silk_rate = limit(silk_rate, 5000, 100000);
    st->targetRate_bps = silk_rate;
    st->useDTX = dtx_enable ? 1 : 0;
    st->useMDIndex = useMDIndex;
    st->variable_HP_smth1_Q15 = 200844;
    st->variable_HP_smth2_Q15 = 200844;
    st->first_frame_after_reset = 1;
    vad_init(&st->vad);
    for (int i = 0; i < 3; i++) st->nsq[i].prev_inv_gain_Q16 = 65536;
    // setup_fs (control_codec_FIX.c:232-317): only the centre NSQ gets lagPrev = 100 (App. A Q5)
    st->prevLag = 100;
    st->prev_sigtype = 1;
    st->pf_lagPrev = 100;
    st->LastGainIndex = 1;
    st->nsq[0].lagPrev = 100;
    // setup_rate (control_codec_FIX.c:319-389)
    const i32* rt = SB_T(target_rate_table_nb);
    const i32* snr = SB_T(snr_table_q1);
    i32 md_rate = silk_rate / 2;
    for (int k = 1; k < 8; k++) {
        if (md_rate < rt[k]) {
            i32 frac_Q6 = shl(md_rate - rt[k - 1], 6) / (rt[k] - rt[k - 1]);
            st->SNRPerMD_dB_Q7 = shl(snr[k - 1], 6) + mulw(frac_Q6, snr[k] - snr[k - 1]);
            break;
        }
    }
    for (int k = 1; k < 8; k++) {
        if (silk_rate <= rt[k]) {
            i32 frac_Q6 = shl(silk_rate - rt[k - 1], 6) / (rt[k] - rt[k - 1]);
            st->SNR_dB_Q7 = shl(snr[k - 1], 6) + mulw(frac_Q6, snr[k] - snr[k - 1]);
            break;
        }
    }
    st->hb_first = 1;
    return 0;
}

// One low-band / high-band output pair of the 64-tap QMF analysis (AGR_BWE_qmf.c:66-79): x = [63 history | N new]
// samples (already >> 1), aa = the 64 coefficients.  Products are summed mod 2^32, so any order gives the same bits.
SB_HD void qmf_output_pair(const i16* x, const i16* aa, int k, i16* lo, i16* hi) {
    enum { M = 64 };
    const int i = 2 * k;
    const i16* x2 = x + M - 1;
    i32 y1k = 0, y2k = 0;
#pragma unroll 4
    for (int j = 0; j < M / 2; j += 2) {
        i32 a0 = aa[M - j - 1], a1 = aa[M - j - 2];
        i32 s0 = (i16)(x[i + j] + x2[i - j]), d0 = (i16)(x[i + j] - x2[i - j]);
        i32 s1 = (i16)(x[i + j + 1] + x2[i - j - 1]), d1 = (i16)(x[i + j + 1] - x2[i - j - 1]);
        y1k = addw(y1k, a0 * s0);
        y2k = subw(y2k, a0 * d0);
        y1k = addw(y1k, a1 * s1);
        y2k = addw(y2k, a1 * d1);
    }
    i32 v1 = addw(y1k, 16384) >> 15, v2 = addw(y2k, 16384) >> 15;
    *lo = (i16)(v1 > 32767 ? 32767 : (v1 < -32767 ? -32767 : v1));
    *hi = (i16)(v2 > 32767 ? 32767 : (v2 < -32767 ? -32767 : v2));
}

// ---- AGR_Sate_qmf_decomp (AGR_BWE_qmf.c:38-80), N = 640 or 320, M = 64, fixed point ------------------------------
SB_FN void qmf_decomp(const i16* xx, i16* y1, i16* y2, i16* mem, int N) {
    enum { M = 64 };
    i16 x[PACKET + M - 1];
    const i16* aa = SB_T(qmf_fix);
    for (int i = 0; i < M - 1; i++) x[i] = mem[M - i - 2];
#ifdef __CUDA_ARCH__
    if ((((size_t)xx) & 15) == 0) {          // PCM rows of the batch buffers: 128-bit loads straight from global memory
        const int4* src = reinterpret_cast<const int4*>(xx);
#pragma unroll 2
        for (int i = 0; i < N / 8; i++) {
            const int4 v = src[i];
            i16* d = &x[8 * i + M - 1];
            d[0] = (i16)((i16)v.x >> 1); d[1] = (i16)(v.x >> 17); d[2] = (i16)((i16)v.y >> 1); d[3] = (i16)(v.y >> 17);
            d[4] = (i16)((i16)v.z >> 1); d[5] = (i16)(v.z >> 17); d[6] = (i16)((i16)v.w >> 1); d[7] = (i16)(v.w >> 17);
        }
    } else
#endif
    for (int i = 0; i < N; i++) x[i + M - 1] = (i16)(xx[i] >> 1);
    for (int i = 0; i < M - 1; i++) mem[i] = x[N + M - 2 - i];   // == xx[N - i - 1] >> 1
    for (int k = 0; k < N / 2; k++) qmf_output_pair(x, aa, k, &y1[k], &y2[k]);
}

// ---- AGR_Sate_lsp_quant_highband (AGR_BWE_quant_highband.c:23-104) ----------------------------------------------
SB_FN i32 hb_lsp_quant(i32* lsp) {
    i32 weight[HB_ORDER];
    nlsf_vq_weights_laroia(weight, lsp, HB_ORDER);
    const i16* cb1 = SB_T(hb_lsp_cb1_fix);
    const i16* cb2 = SB_T(hb_lsp_cb2_fix);
    i32 min_dist = SB_I32_MAX; int idx1 = 0;
    for (int i = 0; i < 256; i++) {
        i32 dist = 0;
        for (int j = 0; j < HB_ORDER; j++) { i32 t = lsp[j] - cb1[i * HB_ORDER + j]; dist = smlabb(dist, t, t); }
        if (dist < min_dist) { min_dist = dist; idx1 = i; }
    }
    for (int i = 0; i < HB_ORDER; i++) lsp[i] -= cb1[idx1 * HB_ORDER + i];
    i32 best = SB_I32_MAX; int idx2 = 0;
    for (int i = 0; i < 16; i++) {
        i32 dist = 0;
        for (int j = 0; j < HB_ORDER; j++) { i32 t = subw(lsp[j], cb2[i * HB_ORDER + j]); dist = smlawb(dist, smulbb(t, t), weight[j]); }
        if (dist < best) { best = dist; idx2 = i; }
    }
    for (int i = 0; i < HB_ORDER; i++) lsp[i] = (i32)cb1[idx1 * HB_ORDER + i] + (i32)cb2[idx2 * HB_ORDER + i];
    return shl(idx2, 8) + idx1;
}

// ---- AGR_Bwe_encode_frame_FIX (AGR_BWE_encode_frame_FIX.c:8-82), split in two because the gain needs the low-band
// excitation produced by the noise-shaping quantiser: (1) buffer update, LPC analysis, LSP VQ, per-sub-frame residual
// energy of the high band; (2) gain = 16*sqrt(E_hb)/sqrt(E_lb_exc), 32-level VQ, bit packing (12 + 4*5 bits, MSB first).
template <int F> SB_FN void hb_analyse_frame_t(EncCore* st, const i16* high, i32* lsp_idx_out, i32* nrg0_out) {
    const int LPCF = 80, SF = F >> 2;   // LPC block and sub-frame lengths; F = frame length (compile-time: the loops unroll)
    for (int i = 0; i < F; i++) st->x_hb_buf[F + 40 + i] = high[i];
    // AGR_Sate_find_HB_LPC_FIX: 4 blocks of (80 + 8) samples, hop 80, starting 8 samples before the frame
    i16 LPC_in_pre[4 * (LPCF + HB_ORDER)];
    {
        const i16* xp = st->x_hb_buf + F - HB_ORDER;
        i16* d = LPC_in_pre;
        for (int k = 0; k < 4; k++) {
            for (int i = 0; i < LPCF + HB_ORDER; i++) d[i] = xp[i];
            d += LPCF + HB_ORDER;
            xp += LPCF;
        }
    }
    i32 NLSF_Q15[HB_ORDER], interp, prev_dummy[HB_ORDER];
    for (int i = 0; i < HB_ORDER; i++) prev_dummy[i] = 0;
    find_lpc(NLSF_Q15, &interp, prev_dummy, 0, HB_ORDER, LPC_in_pre, LPCF + HB_ORDER);
    *lsp_idx_out = hb_lsp_quant(NLSF_Q15);
    i16 coef[HB_ORDER], exc[2 * SUBFR];
    nlsf2a_stable(coef, NLSF_Q15, HB_ORDER);
    const i16* p_hb = st->x_hb_buf + F;
    for (int sub = 0; sub < 4; sub++) {
        lpc_analysis_filter_zero_state(p_hb, coef, exc, SF, HB_ORDER);
        i32 nrg0 = 0;
        for (int i = 0; i < SF; i++) nrg0 = addw(nrg0, (i32)exc[i] * (i32)exc[i]);
        nrg0_out[sub] = sqrt_approx(nrg0);
        p_hb += SF;
    }
    for (int i = 0; i < F + 40; i++) st->x_hb_buf[i] = st->x_hb_buf[F + i];
    st->hb_first = 0;
}
SB_FN void hb_analyse_frame(EncCore* st, const i16* high, i32* lsp_idx_out, i32* nrg0_out) {
    if (st->hb_frame == HB_FRAME) hb_analyse_frame_t<HB_FRAME>(st, high, lsp_idx_out, nrg0_out);
    else hb_analyse_frame_t<2 * HB_FRAME>(st, high, lsp_idx_out, nrg0_out);
}
// r16 = (int16)(low-band excitation Q10 >> 10), the only form in which AGR_BWE_encode_frame_FIX.c:56-58 uses it
// (sf = sub-frame length: 40, or 80 when one high-band frame spans both codec frames of the packet)
SB_FN void hb_pack_frame(i32 lsp_idx, const i32* nrg0, const i16* r16, u8* out4, int sf = SUBFR) {
    i32 gain_idx[4];
    for (int sub = 0; sub < 4; sub++) {
        i32 nrg1 = 0;
        for (int i = 0; i < sf; i++) nrg1 = smlabb(nrg1, r16[sub * sf + i], r16[sub * sf + i]);
        nrg1 = sqrt_approx(nrg1);
        i16 gain = (i16)(shl(nrg0[sub] + 1, 4) / (nrg1 + 1));
        i32 md = SB_I32_MAX; int gi = 0;
        for (int i = 0; i < 32; i++) {
            i16 t = (i16)(gain - SB_T(hb_gain_cb_fix)[i]);
            i32 dist = smulbb(t, t);
            if (dist < md) { md = dist; gi = i; }
        }
        gain_idx[sub] = gi;
    }
    u32 w = ((u32)lsp_idx & 0xFFF) << 20 | ((u32)gain_idx[0] << 15) | ((u32)gain_idx[1] << 10) | ((u32)gain_idx[2] << 5) | (u32)gain_idx[3];
    out4[0] = (u8)(w >> 24); out4[1] = (u8)(w >> 16); out4[2] = (u8)(w >> 8); out4[3] = (u8)w;
}

// Working set of stage A for one stream in the scalar model (tests/hostsim); the device kernels use CoopWork (sb_coop.cuh).
struct EncAnalysisWork {
    i16 low[PACKET / 2], high[PACKET / 2];
    i16 pIn_HP[FRAME];
    i16 res_pitch[2 * FRAME + LA_PITCH];
    EncCtrl c;                 // control of the frame being analysed (copied to EncScratch when the frame is done)
    i16 xfw[FRAME];
    i32 vadFlag;
    const NlsfFastTabs* nlsf_fast;   // optional copies of the NLSF codebooks (null: use the global tables)
};

// SKP_Silk_encode_frame_FIX (encode_frame_FIX.c:34-131, 151-165, 199-208) up to the quantiser, for one frame (scalar model).
SB_FN void encode_frame_analysis(EncCore* st, EncAnalysisWork* W, const i16* pIn, int frame_in_packet) {
    EncCtrl* c = &W->c;
    i16* x_frame = st->x_buf + FRAME;
    c->Seed = st->frameCounter++ & 3;
    vad_get_sa_q8(&st->vad, &st->speech_activity_Q8, c->input_quality_bands_Q15, &c->input_tilt_Q15, pIn);
    hp_variable_cutoff(st, c, W->pIn_HP, pIn);
    for (int i = 0; i < FRAME; i++) x_frame[LA_SHAPE + i] = W->pIn_HP[i];  // LP_variable_cutoff is a copy (transition_frame_no == 0)
    find_pitch_lags(st, c, W->res_pitch, x_frame);
    noise_shape_analysis(st, c, W->res_pitch + FRAME, x_frame);
    prefilter(st, c, W->xfw, x_frame);
    find_pred_coefs(st, c, W->res_pitch, frame_in_packet, W->nlsf_fast);
    process_gains(st, c, frame_in_packet);
    vad_flag_and_dtx(st, &W->vadFlag);
    for (int i = 0; i < FRAME + LA_SHAPE; i++) st->x_buf[i] = st->x_buf[FRAME + i];
    st->prev_sigtype = c->sigtype;
    st->prevLag = c->pitchL[NB_SUBFR - 1];
    st->first_frame_after_reset = 0;
}

// stage A (scalar model): band split, two core frames, high-band frames
SB_FN void enc_packet_analysis(EncCore* st, EncAnalysisWork* W, const i16* pcm, EncScratch* scr, const i16* bands = nullptr) {
    const int nf = st->frames_per_packet;
    if (bands) {
        const int half = nf * FRAME;
        for (int i = 0; i < half; i++) { W->low[i] = bands[i]; W->high[i] = bands[half + i]; }
    } else {
        qmf_decomp(pcm, W->low, W->high, st->qmf_mem, nf * 2 * FRAME);
    }
    for (int f = 0; f < nf; f++) {
        encode_frame_analysis(st, W, W->low + f * FRAME, f);
        scr->c[f] = W->c;
        for (int i = 0; i < FRAME; i++) scr->xfw[f][i] = W->xfw[i];
        scr->vadFlag[f] = W->vadFlag;
    }
    scr->dtx_drop = (st->useDTX && st->inDTX) ? 1 : 0;
    const int nhb = nf * FRAME / st->hb_frame;    // high-band frames per packet
    for (int f = 0; f < nhb; f++) hb_analyse_frame(st, W->high + f * st->hb_frame, &scr->hb_lsp_idx[f], scr->hb_nrg0[f]);
}

// stage C: AGR_Sate_Encoder_Encode tail -- returns the byte count; nBytesOut[0] = total, nBytesOut[1] = len(MD2) + 8
SB_FN i32 enc_packet_finish(EncCore* st, const EncScratch* scr, u8* rcbuf /* MAX_PAYLOAD scratch */, u8* out, i32 out_cap, i16* nBytesOut) {
    int nb[2];
    int ok = 1;
    int written = 0;
    const int nf = st->frames_per_packet, nhb = nf * FRAME / st->hb_frame, hb_bytes = 4 * nhb;
    for (int k = 0; k < 2; k++) {
        RangeEnc rc;
        rc_enc_init(&rc, rcbuf, MAX_PAYLOAD);
        for (int f = 0; f < nf; f++) {
            encode_parameters(&rc, st, &scr->c[f], k, f, scr->vadFlag[f], scr->q_md[f][k]);
            // frame terminator: MORE_FRAMES (1) while frames follow in this packet, LAST_FRAME (0) after the last one
            rc_encode(&rc, f < nf - 1 ? 1 : 0, SB_T(frame_term_cdf));
        }
        rc_get_length(&rc, &nb[k]);
        if (rc.error) ok = 0;
        if (k == 1 && nb[0] + nb[1] > MAX_PAYLOAD) ok = 0;  // pnBytesOut[0] >= nMDBytes (encode_frame_FIX.c:245)
        if (ok) {
            rc_enc_wrap_up(&rc);
            for (int i = 0; i < nb[k]; i++) if (written + i < out_cap) out[written + i] = rcbuf[i];
            written += nb[k];
        }
    }
    if (!ok) { nb[0] = nb[1] = 0; }
    if (scr->dtx_drop) { nb[0] = nb[1] = 0; }
    u8 hb[8];
    for (int f = 0; f < nhb; f++) hb_pack_frame(scr->hb_lsp_idx[f], scr->hb_nrg0[f], &scr->r16[0][0] + f * st->hb_frame, hb + 4 * f, st->hb_frame >> 2);
    int lb = nb[0] + nb[1];
    int total = lb + hb_bytes;
    for (int i = 0; i < hb_bytes; i++) if (lb + i < out_cap) out[lb + i] = hb[i];
    if (lb) { nBytesOut[0] = (i16)total; nBytesOut[1] = (i16)(nb[1] + hb_bytes); }
    else { nBytesOut[0] = 0; nBytesOut[1] = 0; }
    return imin(out_cap, total);
}

// Whole packet on one thread (host model used by tests/hostsim; the device runs stage B as a warp-per-stream kernel).
struct EncPacketWork {
    EncAnalysisWork a;
    EncScratch scr;
    NsqWork nsq;
    i32 r[FRAME];
    u8 rcbuf[MAX_PAYLOAD];
};
// stages B (scalar model of the quantiser) + C on one thread
SB_FN i32 enc_packet_quantise_and_code(EncState* st, EncPacketWork* W, u8* out, i32 out_cap, i16* nBytesOut) {
    for (int f = 0; f < st->frames_per_packet; f++) {
        nsq_del_dec(st, &W->scr.c[f], &W->nsq, W->scr.xfw[f], (i8*)0, W->scr.q_md[f][0], W->scr.q_md[f][1], W->r);
        for (int i = 0; i < FRAME; i++) W->scr.r16[f][i] = (i16)(W->r[i] >> 10);
    }
    return enc_packet_finish(st, &W->scr, W->rcbuf, out, out_cap, nBytesOut);
}
SB_FN i32 enc_packet(EncState* st, EncPacketWork* W, const i16* pcm, u8* out, i32 out_cap, i16* nBytesOut) {
    enc_packet_analysis(st, &W->a, pcm, &W->scr);
    return enc_packet_quantise_and_code(st, W, out, out_cap, nBytesOut);
}

}  // namespace sb
