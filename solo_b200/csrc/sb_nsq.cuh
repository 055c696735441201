// solo_b200 -- multiple-description delayed-decision noise-shaping quantiser (one 20 ms frame).
//
// Three coupled quantisers (centre, description 1, description 2), four delayed-decision states each.
// Reproduces /root/reference/JC1_SDK_SRC_ARM/src/libSATECodec/SKP_Silk_NSQ_del_dec.c:925-1136 (frame),
// :1328-1570 (per-sample loop) with helpers :154-898, :1152-1325, :1572-1692 and Agora_SILK_func.c:7-200.
// Work the reference does but whose results are never observable is dropped:
// the Q_Q10 / X_Q10 / Rd_Q10 / *_md rings, NSQ.q_Q10, the prd[] arrays and SeedInit.  The three
// quantisers advance their ring index, sLTP_buf_idx and sLTP_shp_buf_idx in lock-step in the reference,
// so one copy of each index is kept.  Agora_Silk_RDCx1 recomputes SKP_INVERSE32_varQ(DeltaGains) per
// sample per state (:590); it is a pure function of the sub-frame constants and is hoisted.
#pragma once
#include "sb_sigproc.cuh"
#include "sb_state.cuh"

namespace sb {

// Delayed-decision history.  The reference gives every state private 32-deep rings and memcpy()s all of them when a
// survivor replaces another state (SKP_Silk_copy_del_dec_state, ~1.5 KB x 3 quantisers per copy).  Here each ring
// position holds one entry per *writing* state (slot) and every state carries a 64-bit path word: 2 bits per ring
// position naming the slot that holds its value there.  All states write position `idx` in the same sample, so an entry
// is never referenced after it has been overwritten; a state copy is a copy of the path word.
struct NsqTab {
    i32 RandState[DD_DELAY][N_DD];
    i32 Xq_Q10[DD_DELAY][N_DD];
    i32 Pred_Q16[DD_DELAY][N_DD];
    i32 Shape_Q10[DD_DELAY][N_DD];
    i32 exc_Q10[DD_DELAY][N_DD];
    i8 Q_Q0[DD_DELAY][N_DD];
};
struct NsqDelDec {
    u64 path;
    i32 sAR2_Q14[SHAPE_ORDER];
    i32 sLPC_Q14[SUBFR + DD_DELAY];
    i32 LF_AR_Q12, Seed, Seed2, SeedInit2, RD_Q10;
};
SB_HD int nsq_slot(const NsqDelDec* d, int idx) { return (int)((d->path >> (2 * idx)) & 3); }
struct NsqSample {
    i32 Q_Q0, Q_Q10, RD_Q10, xq_Q14, LF_AR_Q12, sLTP_shp_Q10, LPC_exc_Q16, exc_Q10, Rd_ind_Q10;
};

// SKP_Silk_copy_del_dec_state (:1669-1692)
SB_FN void nsq_copy_state(NsqDelDec* d, const NsqDelDec* s, int lpc_idx) {
    d->path = s->path;
    for (int i = 0; i < SHAPE_ORDER; i++) d->sAR2_Q14[i] = s->sAR2_Q14[i];
    for (int i = 0; i < DD_DELAY; i++) d->sLPC_Q14[lpc_idx + i] = s->sLPC_Q14[lpc_idx + i];
    d->LF_AR_Q12 = s->LF_AR_Q12; d->Seed = s->Seed; d->Seed2 = s->Seed2; d->SeedInit2 = s->SeedInit2; d->RD_Q10 = s->RD_Q10;
}

// Agora_Silk_RDCx1 (:577-667) with the loop-invariant inverse passed in
SB_FN void nsq_rdcx1(const NsqDelDec* dd, NsqSample* ss, i32 r_Q10, i32 r_p_Q10, i32 inv_gain_Q16, i32 Lambda_Q10, i32 offset_Q10) {
    i32 q1, q2, rd1, rd2, rr;
    r_p_Q10 = smulww(inv_gain_Q16, r_p_Q10);
    r_Q10 = subw(r_Q10, offset_Q10);
    r_p_Q10 = subw(r_p_Q10, offset_Q10);
    r_Q10 = limit(r_Q10, -(64 << 10), 64 << 10);
    if (r_Q10 < -1536) {
        q1 = shl(rshift_round(r_Q10, 10), 10);
        rr = subw(r_p_Q10, q1);
        rd1 = smlabb(mulw(negw(addw(q1, offset_Q10)), Lambda_Q10), rr, rr) >> 10;
        q2 = addw(shl(rshift_round(r_Q10, 10), 10), 1024);
        rr = subw(r_p_Q10, q2);
        rd2 = smlabb(mulw(negw(addw(q2, offset_Q10)), Lambda_Q10), rr, rr) >> 10;
    } else if (r_Q10 > 512) {
        q1 = shl(rshift_round(r_Q10, 10), 10);
        rr = subw(r_p_Q10, q1);
        rd1 = smlabb(mulw(addw(q1, offset_Q10), Lambda_Q10), rr, rr) >> 10;
        q2 = subw(shl(rshift_round(r_Q10, 10), 10), 1024);
        rr = subw(r_p_Q10, q2);
        rd2 = smlabb(mulw(addw(q2, offset_Q10), Lambda_Q10), rr, rr) >> 10;
    } else {
        q2 = 0;
        rr = subw(r_p_Q10, q2);
        rd2 = smlabb(mulw(addw(q2, offset_Q10), Lambda_Q10), rr, rr) >> 10;
        q1 = -1024;
        rr = subw(r_p_Q10, q1);
        rd1 = smlabb(mulw(negw(addw(q1, offset_Q10)), Lambda_Q10), rr, rr) >> 10;
    }
    if (rd1 < rd2) {
        ss[0].RD_Q10 = addw(dd->RD_Q10, rd1); ss[1].RD_Q10 = addw(dd->RD_Q10, rd2);
        ss[0].Q_Q0 = (i8)(q1 >> 10); ss[1].Q_Q0 = (i8)(q2 >> 10);
        ss[0].Q_Q10 = q1; ss[1].Q_Q10 = q2;
        ss[0].Rd_ind_Q10 = rd1; ss[1].Rd_ind_Q10 = rd2;
    } else {
        ss[0].RD_Q10 = addw(dd->RD_Q10, rd2); ss[1].RD_Q10 = addw(dd->RD_Q10, rd1);
        ss[0].Q_Q0 = (i8)(q2 >> 10); ss[1].Q_Q0 = (i8)(q1 >> 10);
        ss[0].Q_Q10 = q2; ss[1].Q_Q10 = q1;
        ss[0].Rd_ind_Q10 = rd2; ss[1].Rd_ind_Q10 = rd1;
    }
    ss[0].Q_Q10 = addw(offset_Q10, ss[0].Q_Q10);
    ss[1].Q_Q10 = addw(offset_Q10, ss[1].Q_Q10);
}

// Agora_Silk_CenterRD (:1152-1325).  Composite c pairs side candidates (a(c), b(c)):
//   c0 = (p1[0],p2[0])  c1 = (p1[1],p2[1])  c2 = (p1[0],p2[1])  c3 = (p1[1],p2[0]);
// the reference's twelve memcpy cases are exactly "new p[s] = old p[a|b(winner s)]".
SB_FN void nsq_center_rd(i32 dd_RD_Q10, NsqSample* sc, NsqSample* s1, NsqSample* s2, i32 res_Q10, i32 Lambda_Q10, i32 offset_Q10) {
    const i32 JL = 90000;  // INTERNAL_JOINT_LAMBDA
    i32 qx[4], rdx[4];
    qx[0] = addw(s1[0].Q_Q10, s2[0].Q_Q10);
    qx[1] = addw(s1[1].Q_Q10, s2[1].Q_Q10);
    qx[2] = addw(s1[0].Q_Q10, s2[1].Q_Q10);
    qx[3] = addw(s1[1].Q_Q10, s2[0].Q_Q10);
    i32 r_temp = subw(res_Q10, offset_Q10);
    for (int s = 0; s < 4; s++) {
        i32 rr = subw(r_temp, qx[s]);
        i32 t = addw(qx[s], offset_Q10);
        if (qx[s] < 0) t = negw(t);
        rdx[s] = smlabb(mulw(t, Lambda_Q10), rr, rr) >> 10;
    }
    rdx[0] = addw(addw(rdx[0], smulww(JL, s1[0].Rd_ind_Q10)), smulww(JL, s2[0].Rd_ind_Q10));
    rdx[1] = addw(addw(rdx[1], smulww(JL, s1[1].Rd_ind_Q10)), smulww(JL, s2[1].Rd_ind_Q10));
    rdx[2] = addw(addw(rdx[2], smulww(JL, s1[0].Rd_ind_Q10)), smulww(JL, s2[1].Rd_ind_Q10));
    rdx[3] = addw(addw(rdx[3], smulww(JL, s1[1].Rd_ind_Q10)), smulww(JL, s2[0].Rd_ind_Q10));
    i32 mn = rdx[0]; int w1 = 0, w2;
    for (int s = 1; s < 4; s++) if (rdx[s] < mn) { mn = rdx[s]; w1 = s; }
    if (w1 == 0) {
        mn = rdx[1]; w2 = 1;
        for (int s = 2; s < 4; s++) if (rdx[s] < mn) { mn = rdx[s]; w2 = s; }
    } else {
        mn = rdx[0]; w2 = 0;
        for (int s = 1; s < 4; s++) if (rdx[s] < mn && s != w1) { mn = rdx[s]; w2 = s; }
    }
    sc[0].RD_Q10 = addw(dd_RD_Q10, rdx[w1]); sc[1].RD_Q10 = addw(dd_RD_Q10, rdx[w2]);
    sc[0].Q_Q0 = qx[w1] >> 10; sc[1].Q_Q0 = qx[w2] >> 10;
    sc[0].Q_Q10 = qx[w1]; sc[1].Q_Q10 = qx[w2];
    sc[0].Rd_ind_Q10 = rdx[w1]; sc[1].Rd_ind_Q10 = rdx[w2];
    const int A = 0xA;  // a(c): bit c = {0,1,0,1}
    const int B = 0x6;  // b(c): bit c = {0,1,1,0}
    NsqSample o1[2] = {s1[0], s1[1]}, o2[2] = {s2[0], s2[1]};
    s1[0] = o1[(A >> w1) & 1]; s1[1] = o1[(A >> w2) & 1];
    s2[0] = o2[(B >> w1) & 1]; s2[1] = o2[(B >> w2) & 1];
}

// Agora_Silk_UnDither (:559-572)
SB_HD void nsq_undither(const NsqDelDec* dd, NsqSample* ss) {
    i32 dither = dd->Seed2 >> 31;
    ss[0].Q_Q10 = subw(ss[0].Q_Q10 ^ dither, dither);
    ss[1].Q_Q10 = subw(ss[1].Q_Q10 ^ dither, dither);
    ss[0].exc_Q10 = ss[0].Q_Q10;
    ss[1].exc_Q10 = ss[1].Q_Q10;
}
// Agora_Silk_UndoPred_And_Shap (:491-530)
SB_HD void nsq_undo_pred(NsqSample* ss, i32 LTP_pred_Q14, i32 LPC_pred_Q10, i32 n_AR_Q10, i32 n_LF_Q10) {
    for (int s = 0; s < 2; s++) {
        i32 LPC_exc_Q10 = addw(ss[s].Q_Q10, rshift_round(LTP_pred_Q14, 4));
        i32 xq_Q10 = addw(LPC_exc_Q10, LPC_pred_Q10);
        i32 sLF_AR_shp_Q10 = subw(xq_Q10, n_AR_Q10);
        ss[s].sLTP_shp_Q10 = subw(sLF_AR_shp_Q10, n_LF_Q10);
        ss[s].LF_AR_Q12 = shl(sLF_AR_shp_Q10, 2);
        ss[s].xq_Q14 = shl(xq_Q10, 4);
        ss[s].LPC_exc_Q16 = shl(LPC_exc_Q10, 6);
    }
}

struct NsqWork {
    NsqDelDec dd[3][N_DD];
    NsqTab tab[3];
    NsqSample ss[3][N_DD][2];
    i32 sLTP_Q16[3][2 * FRAME];
    i16 sLTP[3][2 * FRAME];
    i32 Gain_Q16[DD_DELAY];  // identical in every state of every quantiser (App. A Q6)
    i32 x_sc_Q10[SUBFR];
};

// flush `n` delayed samples of state `w` of quantiser `qz` to the outputs (shared by the k==2 reset and the frame end)
SB_FN void nsq_flush(NsqState* ns, NsqWork* W, int qz, int w, int smpl_buf_idx, int decisionDelay, int sig_off, int shp_idx,
                     int ltp_idx, i8* q, i32* r, int write_pred) {
    const NsqDelDec* psDD = &W->dd[qz][w];
    const NsqTab* T = &W->tab[qz];
    int last = smpl_buf_idx + decisionDelay;
    for (int i = 0; i < decisionDelay; i++) {
        last = (last - 1) & DD_MASK;
        const int sl = nsq_slot(psDD, last);
        int o = sig_off + i - decisionDelay;
        if (q) q[o] = T->Q_Q0[last][sl];
        if (r) r[o] = T->exc_Q10[last][sl];
        ns->xq[FRAME + o] = (i16)sat16(rshift_round(smulww(T->Xq_Q10[last][sl], W->Gain_Q16[last]), 10));
        ns->sLTP_shp_Q10[shp_idx - decisionDelay + i] = T->Shape_Q10[last][sl];
        if (write_pred) W->sLTP_Q16[qz][ltp_idx - decisionDelay + i] = T->Pred_Q16[last][sl];
    }
}

// SKP_Silk_NSQ_del_dec (:925-1136).  x = prefiltered input (160), q_md[2] = pulses of the two descriptions,
// q_c = centre pulses (may be NULL), r = centre excitation Q10 (feeds the high-band gain).
SB_FN void nsq_del_dec(EncState* st, EncCtrl* c, NsqWork* W, const i16* x, i8* q_c, i8* q_md0, i8* q_md1, i32* r) {
    NsqState* NS[3] = {&st->nsq[0], &st->nsq[1], &st->nsq[2]};
    i8* Q[3] = {q_c, q_md0, q_md1};
    i32 lag[3] = {NS[0]->lagPrev, NS[1]->lagPrev, NS[2]->lagPrev};
    const i32 offset_Q10 = SB_T(quant_offsets_q10)[c->sigtype * 2 + c->QuantOffsetType];
    int smpl_buf_idx = 0;
    int decisionDelay = imin(DD_DELAY, SUBFR);
    if (c->sigtype == 0) {
        for (int k = 0; k < NB_SUBFR; k++) decisionDelay = imin(decisionDelay, c->pitchL[k] - LTP_ORDER / 2 - 1);
    } else if (lag[0] > 0) {
        decisionDelay = imin(decisionDelay, lag[0] - LTP_ORDER / 2 - 1);
    }
    const int LSF_flag = c->NLSFInterpCoef_Q2 == 4 ? 0 : 1;
    // Agora_Silk_Init_DelDecState (:154-182)
    for (int qz = 0; qz < 3; qz++) {
        for (int k = 0; k < N_DD; k++) {
            NsqDelDec* d = &W->dd[qz][k];
            memset(d, 0, sizeof(NsqDelDec));
            d->path = 0x5555555555555555ull * (u64)k;  // every ring position -> own slot
            d->Seed = (k + c->Seed) & 3;
            d->Seed2 = d->Seed;
            d->SeedInit2 = d->Seed;
            d->LF_AR_Q12 = NS[qz]->sLF_AR_shp_Q12;
            for (int i = 0; i < DD_DELAY; i++) d->sLPC_Q14[i] = NS[qz]->sLPC_Q14[i];
            for (int i = 0; i < SHAPE_ORDER; i++) d->sAR2_Q14[i] = NS[qz]->sAR2_Q14[i];
        }
        memset(&W->tab[qz], 0, sizeof(NsqTab));
        for (int k = 0; k < N_DD; k++) W->tab[qz].Shape_Q10[0][k] = NS[qz]->sLTP_shp_Q10[FRAME - 1];
        for (int i = 0; i < 2 * FRAME; i++) { W->sLTP_Q16[qz][i] = 0; W->sLTP[qz][i] = 0; }
    }
    for (int i = 0; i < DD_DELAY; i++) W->Gain_Q16[i] = 0;
    int shp_idx = FRAME, ltp_idx = FRAME;  // sLTP_shp_buf_idx / sLTP_buf_idx
    int subfr = 0;

    for (int k = 0; k < NB_SUBFR; k++) {
        const i16* A_Q12 = c->PredCoef_Q12[(k >> 1) | (1 - LSF_flag)];
        const i16* B_Q14 = &c->LTPCoef_Q14[k * LTP_ORDER];
        const i16* AR_shp_Q13 = &c->AR2_Q13[k * SHAPE_ORDER];
        i32 HarmShapeFIRPacked_Q14 = c->HarmShapeGain_Q14[k] >> 2;
        HarmShapeFIRPacked_Q14 |= shl(c->HarmShapeGain_Q14[k] >> 1, 16);
        const int sig_off = k * SUBFR;
        int rewhite = 0;
        if (c->sigtype == 0) {
            lag[0] = lag[1] = lag[2] = c->pitchL[k];
            if ((k & (3 - shl(LSF_flag, 1))) == 0) {
                if (k == 2) subfr = 0;
                // Agora_Silk_DelDec_Rewhitening{,_Side} (:316-486)
                int Winner_ind = 0;
                if (k == 2) {
                    i32 RDmin = W->dd[0][0].RD_Q10;
                    for (int i = 1; i < N_DD; i++) if (W->dd[0][i].RD_Q10 < RDmin) { RDmin = W->dd[0][i].RD_Q10; Winner_ind = i; }
                }
                for (int qz = 0; qz < 3; qz++) {
                    if (k == 2) {
                        for (int i = 0; i < N_DD; i++) if (i != Winner_ind) W->dd[qz][i].RD_Q10 = addw(W->dd[qz][i].RD_Q10, SB_I32_MAX >> 4);
                        nsq_flush(NS[qz], W, qz, Winner_ind, smpl_buf_idx, decisionDelay, sig_off, shp_idx, ltp_idx, Q[qz], qz == 0 ? r : (i32*)0, 0);
                    }
                    int start_idx = FRAME - lag[qz] - LPC_ORDER - LTP_ORDER / 2;
                    ma_prediction_zero_state(&NS[qz]->xq[start_idx + k * SUBFR], A_Q12, &W->sLTP[qz][start_idx], FRAME - start_idx, LPC_ORDER);
                }
                ltp_idx = FRAME;
                rewhite = 1;
            }
        }
        // Agora_Silk_DelDecScale (:1652-1667) + SKP_Silk_nsq_del_dec_scale_states (:1572-1647)
        i32 inv_gain_Q16 = inverse32_varq(imax(c->Gains_Q16[k], 1), 32);
        inv_gain_Q16 = imin(inv_gain_Q16, 32767);
        for (int i = 0; i < SUBFR; i++) W->x_sc_Q10[i] = smulbb(x[sig_off + i], inv_gain_Q16) >> 6;
        for (int qz = 0; qz < 3; qz++) {
            NsqState* ns = NS[qz];
            const int lagk = c->pitchL[k];
            if (rewhite) {
                i32 inv_gain_Q32 = shl(inv_gain_Q16, 16);
                if (k == 0) inv_gain_Q32 = shl(smulwb(inv_gain_Q32, c->LTP_scale_Q14), 2);
                for (int i = ltp_idx - lagk - LTP_ORDER / 2; i < ltp_idx; i++) W->sLTP_Q16[qz][i] = smulwb(inv_gain_Q32, W->sLTP[qz][i]);
            }
            if (inv_gain_Q16 != ns->prev_inv_gain_Q16) {
                i32 gain_adj_Q16 = div32_varq(inv_gain_Q16, ns->prev_inv_gain_Q16, 16);
                for (int i = shp_idx - SUBFR * NB_SUBFR; i < shp_idx; i++) ns->sLTP_shp_Q10[i] = smulww(gain_adj_Q16, ns->sLTP_shp_Q10[i]);
                if (!rewhite)
                    for (int i = ltp_idx - lagk - LTP_ORDER / 2; i < ltp_idx; i++) W->sLTP_Q16[qz][i] = smulww(gain_adj_Q16, W->sLTP_Q16[qz][i]);
                for (int s = 0; s < N_DD; s++) {
                    NsqDelDec* d = &W->dd[qz][s];
                    d->LF_AR_Q12 = smulww(gain_adj_Q16, d->LF_AR_Q12);
                    for (int i = 0; i < DD_DELAY; i++) d->sLPC_Q14[i] = smulww(gain_adj_Q16, d->sLPC_Q14[i]);
                    for (int i = 0; i < SHAPE_ORDER; i++) d->sAR2_Q14[i] = smulww(gain_adj_Q16, d->sAR2_Q14[i]);
                }
                for (int i = 0; i < DD_DELAY; i++)
                    for (int s = 0; s < N_DD; s++) {
                        W->tab[qz].Pred_Q16[i][s] = smulww(gain_adj_Q16, W->tab[qz].Pred_Q16[i][s]);
                        W->tab[qz].Shape_Q10[i][s] = smulww(gain_adj_Q16, W->tab[qz].Shape_Q10[i][s]);
                    }
            }
            ns->prev_inv_gain_Q16 = inv_gain_Q16;
        }

        // ---- SKP_Silk_md_noise_shape_quantizer_del_dec (:1328-1570) ----
        const i32 Gain_Q16 = c->Gains_Q16[k];
        const i32 Tilt_Q14 = c->Tilt_Q14[k], LF_shp_Q14 = c->LF_shp_Q14[k], Lambda_Q10 = c->Lambda_Q10;
        i32 invg = inverse32_varq(imax(c->DeltaGains_Q16, 1), 32);
        const i32 inv_gain_p1 = invg, inv_gain_p2 = 65536 - invg;
        const i32 DeltaGains_p1 = inverse32_varq(imax(inv_gain_p1, 1), 32);
        const i32 DeltaGains_p2 = inverse32_varq(imax(inv_gain_p2, 1), 32);
        const i32 offset_p1 = smulww(inv_gain_p1, offset_Q10), offset_p2 = smulww(inv_gain_p2, offset_Q10);
        // role of the two descriptions alternates with the sub-frame parity (:1497-1534)
        const int swap = (subfr % 2) >= 1;
        const i32 ig[3] = {0, swap ? inv_gain_p2 : inv_gain_p1, swap ? inv_gain_p1 : inv_gain_p2};
        const i32 dg[3] = {0, swap ? DeltaGains_p2 : DeltaGains_p1, swap ? DeltaGains_p1 : DeltaGains_p2};
        const i32 of[3] = {0, swap ? offset_p2 : offset_p1, swap ? offset_p1 : offset_p2};
        const i32 rdc_inv[3] = {0, inverse32_varq(imax(dg[1], 1), 32), inverse32_varq(imax(dg[2], 1), 32)};
        int shp_lag[3], pred_lag[3];
        for (int qz = 0; qz < 3; qz++) { shp_lag[qz] = shp_idx - lag[qz] + 1; pred_lag[qz] = ltp_idx - lag[qz] + LTP_ORDER / 2; }

        for (int i = 0; i < SUBFR; i++) {
            i32 LTP_pred_Q14[3], n_LTP_Q14[3];
            for (int qz = 0; qz < 3; qz++) {
                i32 p = 0;
                if (c->sigtype == 0) {
                    const i32* pl = &W->sLTP_Q16[qz][pred_lag[qz]];
                    for (int j = 0; j < LTP_ORDER; j++) p = smlawb(p, pl[-j], B_Q14[j]);
                    pred_lag[qz]++;
                }
                LTP_pred_Q14[qz] = p;
                i32 n = 0;
                if (lag[0] > 0) {
                    const i32* sl = &NS[qz]->sLTP_shp_Q10[shp_lag[qz]];
                    n = smulwb(addw(sl[0], sl[-2]), HarmShapeFIRPacked_Q14);
                    n = smlawt(n, sl[-1], HarmShapeFIRPacked_Q14);
                    n = shl(n, 6);
                    shp_lag[qz]++;
                }
                n_LTP_Q14[qz] = n;
            }
            for (int s = 0; s < N_DD; s++) {
                i32 LPC_pred_Q10[3], n_AR_Q10[3], n_LF_Q10[3], r_Q10[3];
                for (int qz = 0; qz < 3; qz++) {
                    NsqDelDec* d = &W->dd[qz][s];
                    const i32* ps = &d->sLPC_Q14[DD_DELAY - 1 + i];
                    i32 lp = 0;
                    for (int j = 0; j < LPC_ORDER; j++) lp = smlawb(lp, ps[-j], A_Q12[j]);
                    LPC_pred_Q10[qz] = lp;
                    // Agora_Silk_STS (Agora_SILK_func.c:85-127): warped all-pass chain
                    i32 tmp2 = smlawb(ps[0], d->sAR2_Q14[0], WARPING_Q16);
                    i32 tmp1 = smlawb(d->sAR2_Q14[0], subw(d->sAR2_Q14[1], tmp2), WARPING_Q16);
                    d->sAR2_Q14[0] = tmp2;
                    i32 n_AR = smulwb(tmp2, AR_shp_Q13[0]);
                    for (int j = 2; j < SHAPE_ORDER; j += 2) {
                        tmp2 = smlawb(d->sAR2_Q14[j - 1], subw(d->sAR2_Q14[j], tmp1), WARPING_Q16);
                        d->sAR2_Q14[j - 1] = tmp1;
                        n_AR = smlawb(n_AR, tmp1, AR_shp_Q13[j - 1]);
                        tmp1 = smlawb(d->sAR2_Q14[j], subw(d->sAR2_Q14[j + 1], tmp2), WARPING_Q16);
                        d->sAR2_Q14[j] = tmp2;
                        n_AR = smlawb(n_AR, tmp2, AR_shp_Q13[j]);
                    }
                    d->sAR2_Q14[SHAPE_ORDER - 1] = tmp1;
                    n_AR = smlawb(n_AR, tmp1, AR_shp_Q13[SHAPE_ORDER - 1]);
                    n_AR = n_AR >> 1;
                    n_AR = smlawb(n_AR, d->LF_AR_Q12, Tilt_Q14);
                    n_AR_Q10[qz] = n_AR;
                    // Agora_Silk_LFS (:129-141)
                    i32 n_LF = shl(smulwb(W->tab[qz].Shape_Q10[smpl_buf_idx][nsq_slot(d, smpl_buf_idx)], LF_shp_Q14), 2);
                    n_LF = smlawt(n_LF, d->LF_AR_Q12, LF_shp_Q14);
                    n_LF_Q10[qz] = n_LF;
                    // Agora_Silk_DoPred_And_Shap (:143-163)
                    i32 t = subw(LTP_pred_Q14[qz], n_LTP_Q14[qz]);
                    t = t >> 4;
                    t = addw(t, lp);
                    t = subw(t, n_AR);
                    t = subw(t, n_LF);
                    i32 rq = subw(W->x_sc_Q10[i], t);
                    // Agora_Silk_Dither (:536-552)
                    d->Seed2 = lcg_rand(d->Seed2);
                    d->Seed = lcg_rand(d->Seed);
                    i32 dither = d->Seed2 >> 31;
                    r_Q10[qz] = subw(rq ^ dither, dither);
                }
                i32 r_md1 = smulww(ig[1], r_Q10[0]);
                i32 r_md2 = smulww(ig[2], r_Q10[0]);
                nsq_rdcx1(&W->dd[1][s], W->ss[1][s], r_md1, r_Q10[1], rdc_inv[1], Lambda_Q10, of[1]);
                nsq_rdcx1(&W->dd[2][s], W->ss[2][s], r_md2, r_Q10[2], rdc_inv[2], Lambda_Q10, of[2]);
                nsq_center_rd(W->dd[0][s].RD_Q10, W->ss[0][s], W->ss[1][s], W->ss[2][s], r_Q10[0], Lambda_Q10, offset_p1 + offset_p2);
                nsq_undither(&W->dd[1][s], W->ss[1][s]);
                nsq_undither(&W->dd[2][s], W->ss[2][s]);
                for (int qz = 1; qz < 3; qz++) {
                    W->ss[qz][s][0].Q_Q10 = smulww(dg[qz], W->ss[qz][s][0].Q_Q10);
                    W->ss[qz][s][1].Q_Q10 = smulww(dg[qz], W->ss[qz][s][1].Q_Q10);
                }
                nsq_undither(&W->dd[0][s], W->ss[0][s]);
                for (int qz = 0; qz < 3; qz++) nsq_undo_pred(W->ss[qz][s], LTP_pred_Q14[qz], LPC_pred_Q10[qz], n_AR_Q10[qz], n_LF_Q10[qz]);
            }
            smpl_buf_idx = (smpl_buf_idx - 1) & DD_MASK;
            const int last_smple_idx = (smpl_buf_idx + decisionDelay) & DD_MASK;

            // Agora_Silk_JudgeWinner (:673-752)
            {
                const i32 JL = 90000;
                int RandSyncCtl = 0, Winner_ind = 0;
                i32 RDmin = addw(addw(W->ss[0][0][0].RD_Q10, smulww(W->ss[1][0][0].RD_Q10, JL)), smulww(W->ss[2][0][0].RD_Q10, JL));
                for (int s = 1; s < N_DD; s++) {
                    i32 j = addw(addw(W->ss[0][s][0].RD_Q10, smulww(W->ss[1][s][0].RD_Q10, JL)), smulww(W->ss[2][s][0].RD_Q10, JL));
                    if (j < RDmin) { RDmin = j; Winner_ind = s; }
                }
                i32 wr[3];
                for (int qz = 0; qz < 3; qz++) wr[qz] = W->tab[qz].RandState[last_smple_idx][nsq_slot(&W->dd[qz][Winner_ind], last_smple_idx)];
                for (int s = 0; s < N_DD; s++) {
                    int mism = 0;
                    for (int qz = 0; qz < 3; qz++) mism |= (W->tab[qz].RandState[last_smple_idx][nsq_slot(&W->dd[qz][s], last_smple_idx)] != wr[qz]);
                    if (mism) {
                        RandSyncCtl++;
                        W->ss[0][s][0].RD_Q10 = addw(W->ss[0][s][0].RD_Q10, SB_I32_MAX >> 4);
                        W->ss[0][s][1].RD_Q10 = addw(W->ss[0][s][1].RD_Q10, SB_I32_MAX >> 4);
                    }
                }
                do {
                    i32 RDmax = W->ss[0][0][0].RD_Q10, RDmin2 = W->ss[0][0][1].RD_Q10;
                    int RDmax_ind = 0, RDmin_ind = 0;
                    for (int s = 1; s < N_DD; s++) {
                        if (W->ss[0][s][0].RD_Q10 > RDmax) { RDmax = W->ss[0][s][0].RD_Q10; RDmax_ind = s; }
                        if (W->ss[0][s][1].RD_Q10 < RDmin2) { RDmin2 = W->ss[0][s][1].RD_Q10; RDmin_ind = s; }
                    }
                    if (RDmin2 < RDmax) {
                        for (int qz = 0; qz < 3; qz++) {
                            nsq_copy_state(&W->dd[qz][RDmax_ind], &W->dd[qz][RDmin_ind], i);
                            W->ss[qz][RDmax_ind][0] = W->ss[qz][RDmin_ind][1];
                        }
                    }
                } while (--RandSyncCtl > 0);
            }
            // Agora_Silk_GetWinner (:759-812) / _Side (:817-858)
            {
                const i32 JL = 90000;
                int Winner_ind = 0;
                i32 RDmin = addw(addw(W->ss[0][0][0].RD_Q10, smulww(W->ss[1][0][0].RD_Q10, JL)), smulww(W->ss[2][0][0].RD_Q10, JL));
                for (int s = 1; s < N_DD; s++) {
                    i32 j = addw(addw(W->ss[0][s][0].RD_Q10, smulww(W->ss[1][s][0].RD_Q10, JL)), smulww(W->ss[2][s][0].RD_Q10, JL));
                    if (j < RDmin) { RDmin = j; Winner_ind = s; }
                }
                if (subfr > 0 || i >= decisionDelay) {
                    for (int qz = 0; qz < 3; qz++) {
                        const NsqDelDec* d = &W->dd[qz][Winner_ind];
                        const NsqTab* T = &W->tab[qz];
                        const int sl = nsq_slot(d, last_smple_idx);
                        const int o = sig_off + i - decisionDelay;
                        if (Q[qz]) Q[qz][o] = T->Q_Q0[last_smple_idx][sl];
                        if (qz == 0) r[o] = T->exc_Q10[last_smple_idx][sl];
                        NS[qz]->xq[FRAME + o] = (i16)sat16(rshift_round(smulww(T->Xq_Q10[last_smple_idx][sl], W->Gain_Q16[last_smple_idx]), 10));
                        NS[qz]->sLTP_shp_Q10[shp_idx - decisionDelay] = T->Shape_Q10[last_smple_idx][sl];
                        W->sLTP_Q16[qz][ltp_idx - decisionDelay] = T->Pred_Q16[last_smple_idx][sl];
                    }
                }
                shp_idx++;
                ltp_idx++;
            }
            // Agora_Silk_Update_DelDecState (:863-898)
            for (int qz = 0; qz < 3; qz++) {
                for (int s = 0; s < N_DD; s++) {
                    NsqDelDec* d = &W->dd[qz][s];
                    const NsqSample* p = &W->ss[qz][s][0];
                    d->LF_AR_Q12 = p->LF_AR_Q12;
                    d->sLPC_Q14[DD_DELAY + i] = p->xq_Q14;
                    NsqTab* T = &W->tab[qz];
                    T->Xq_Q10[smpl_buf_idx][s] = p->xq_Q14 >> 4;
                    T->Q_Q0[smpl_buf_idx][s] = (i8)p->Q_Q0;
                    T->Pred_Q16[smpl_buf_idx][s] = p->LPC_exc_Q16;
                    T->Shape_Q10[smpl_buf_idx][s] = p->sLTP_shp_Q10;
                    d->Seed = addw(d->Seed, p->Q_Q0);
                    T->RandState[smpl_buf_idx][s] = d->Seed;
                    d->RD_Q10 = p->RD_Q10;
                    T->exc_Q10[smpl_buf_idx][s] = p->exc_Q10;
                    d->path = (d->path & ~((u64)3 << (2 * smpl_buf_idx))) | ((u64)s << (2 * smpl_buf_idx));
                }
            }
            W->Gain_Q16[smpl_buf_idx] = Gain_Q16;
        }
        // Agora_Silk_Update_DelDecLPCState (:903-920)
        for (int qz = 0; qz < 3; qz++)
            for (int s = 0; s < N_DD; s++)
                for (int i = 0; i < DD_DELAY; i++) W->dd[qz][s].sLPC_Q14[i] = W->dd[qz][s].sLPC_Q14[SUBFR + i];
        subfr++;
    }

    // Agora_Silk_DelDec_UpdateState_And_Output{,_Side} (:184-310)
    int Winner_ind = 0;
    {
        i32 RDmin = W->dd[0][0].RD_Q10;
        for (int s = 1; s < N_DD; s++) if (W->dd[0][s].RD_Q10 < RDmin) { RDmin = W->dd[0][s].RD_Q10; Winner_ind = s; }
    }
    c->Seed = W->dd[0][Winner_ind].SeedInit2;
    for (int qz = 0; qz < 3; qz++) {
        NsqState* ns = NS[qz];
        nsq_flush(ns, W, qz, Winner_ind, smpl_buf_idx, decisionDelay, FRAME, shp_idx, ltp_idx, Q[qz], qz == 0 ? r : (i32*)0, 1);
        const NsqDelDec* d = &W->dd[qz][Winner_ind];
        for (int i = 0; i < DD_DELAY; i++) ns->sLPC_Q14[i] = d->sLPC_Q14[SUBFR + i];
        for (int i = 0; i < SHAPE_ORDER; i++) ns->sAR2_Q14[i] = d->sAR2_Q14[i];
        ns->sLF_AR_shp_Q12 = d->LF_AR_Q12;
        ns->lagPrev = c->pitchL[NB_SUBFR - 1];
        for (int i = 0; i < FRAME; i++) { ns->xq[i] = ns->xq[FRAME + i]; ns->sLTP_shp_Q10[i] = ns->sLTP_shp_Q10[FRAME + i]; }
    }
}

}  // namespace sb
