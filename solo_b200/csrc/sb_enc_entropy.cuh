// solo_b200 -- side-information and pulse entropy coding of one 20 ms frame of one description.
// Reference: /root/reference/JC1_SDK_SRC_ARM/src/libSATECodec/SKP_Silk_encode_parameters.c:33-182,
// SKP_Silk_encode_pulses.c:55-195, SKP_Silk_shell_coder.c:82-120, SKP_Silk_code_signs.c:37-61.
// The centre description is never range-coded: its bytes are discarded by the reference
// (SKP_Silk_enc_API.c:267-270) and have no observable feedback.
#pragma once
#include "sb_rangecoder.cuh"
#include "sb_tables.cuh"

namespace sb {

SB_HD void shell_split(RangeEnc* rc, int p_child1, int p, const u16* shell_table) {
    if (p > 0) rc_encode(rc, p_child1, &shell_table[SB_T(shell_table_offsets)[p]]);
}
// SKP_Silk_shell_encoder (shell_coder.c:82-120)
SB_FN void shell_encoder(RangeEnc* rc, const i32* p0) {
    i32 p1[8], p2[4], p3[2], p4;
    for (int k = 0; k < 8; k++) p1[k] = p0[2 * k] + p0[2 * k + 1];
    for (int k = 0; k < 4; k++) p2[k] = p1[2 * k] + p1[2 * k + 1];
    for (int k = 0; k < 2; k++) p3[k] = p2[2 * k] + p2[2 * k + 1];
    p4 = p3[0] + p3[1];
    shell_split(rc, p3[0], p4, SB_T(shell_table3));
    shell_split(rc, p2[0], p3[0], SB_T(shell_table2));
    shell_split(rc, p1[0], p2[0], SB_T(shell_table1));
    shell_split(rc, p0[0], p1[0], SB_T(shell_table0));
    shell_split(rc, p0[2], p1[1], SB_T(shell_table0));
    shell_split(rc, p1[2], p2[1], SB_T(shell_table1));
    shell_split(rc, p0[4], p1[2], SB_T(shell_table0));
    shell_split(rc, p0[6], p1[3], SB_T(shell_table0));
    shell_split(rc, p2[2], p3[1], SB_T(shell_table2));
    shell_split(rc, p1[4], p2[2], SB_T(shell_table1));
    shell_split(rc, p0[8], p1[4], SB_T(shell_table0));
    shell_split(rc, p0[10], p1[5], SB_T(shell_table0));
    shell_split(rc, p1[6], p2[3], SB_T(shell_table1));
    shell_split(rc, p0[12], p1[6], SB_T(shell_table0));
    shell_split(rc, p0[14], p1[7], SB_T(shell_table0));
}

SB_HD int combine_and_check(i32* out, const i32* in, int max_pulses, int len) {
    for (int k = 0; k < len; k++) {
        i32 sum = in[2 * k] + in[2 * k + 1];
        if (sum > max_pulses) return 1;
        out[k] = sum;
    }
    return 0;
}

// SKP_Silk_encode_pulses (encode_pulses.c:55-195), frame length 160
SB_FN void encode_pulses(RangeEnc* rc, int sigtype, int QuantOffsetType, const i8* q) {
    enum { ITER = FRAME / 16 };
    i32 abs_pulses[FRAME], sum_pulses[ITER], nRshifts[ITER], pulses_comb[8];
    for (int i = 0; i < 8; i++) pulses_comb[i] = 0;
    for (int i = 0; i < FRAME; i++) abs_pulses[i] = q[i] > 0 ? q[i] : -q[i];
    i32* ap = abs_pulses;
    for (int i = 0; i < ITER; i++) {
        nRshifts[i] = 0;
        while (1) {
            int scale_down = combine_and_check(pulses_comb, ap, SB_T(max_pulses_table)[0], 8);
            scale_down += combine_and_check(pulses_comb, pulses_comb, SB_T(max_pulses_table)[1], 4);
            scale_down += combine_and_check(pulses_comb, pulses_comb, SB_T(max_pulses_table)[2], 2);
            sum_pulses[i] = pulses_comb[0] + pulses_comb[1];
            if (sum_pulses[i] > SB_T(max_pulses_table)[3]) scale_down++;
            if (scale_down) {
                nRshifts[i]++;
                for (int k = 0; k < 16; k++) ap[k] = ap[k] >> 1;
            } else break;
        }
        ap += 16;
    }
    i32 minSumBits_Q6 = SB_I32_MAX;
    int RateLevelIndex = 0;
    for (int k = 0; k < 9; k++) {
        const i16* nBits = &SB_T(pulses_per_block_bits_q6)[k * 20];
        i32 sumBits_Q6 = SB_T(rate_levels_bits_q6)[sigtype * 9 + k];
        for (int i = 0; i < ITER; i++) sumBits_Q6 += nRshifts[i] > 0 ? nBits[18 + 1] : nBits[sum_pulses[i]];
        if (sumBits_Q6 < minSumBits_Q6) { minSumBits_Q6 = sumBits_Q6; RateLevelIndex = k; }
    }
    rc_encode(rc, RateLevelIndex, &SB_T(rate_levels_cdf)[sigtype * 10]);
    const u16* cdf_ptr = &SB_T(pulses_per_block_cdf)[RateLevelIndex * 21];
    const u16* cdf_last = &SB_T(pulses_per_block_cdf)[9 * 21];
    for (int i = 0; i < ITER; i++) {
        if (nRshifts[i] == 0) rc_encode(rc, sum_pulses[i], cdf_ptr);
        else {
            rc_encode(rc, 18 + 1, cdf_ptr);
            for (int k = 0; k < nRshifts[i] - 1; k++) rc_encode(rc, 18 + 1, cdf_last);
            rc_encode(rc, sum_pulses[i], cdf_last);
        }
    }
    for (int i = 0; i < ITER; i++) if (sum_pulses[i] > 0) shell_encoder(rc, &abs_pulses[i * 16]);
    for (int i = 0; i < ITER; i++) {
        if (nRshifts[i] > 0) {
            const i8* pp = &q[i * 16];
            int nLS = nRshifts[i] - 1;
            for (int k = 0; k < 16; k++) {
                i32 abs_q = (i8)(pp[k] > 0 ? pp[k] : -pp[k]);
                for (int j = nLS; j > 0; j--) rc_encode(rc, (abs_q >> j) & 1, SB_T(lsb_cdf));
                rc_encode(rc, abs_q & 1, SB_T(lsb_cdf));
            }
        }
    }
    // SKP_Silk_encode_signs (code_signs.c:37-61)
    {
        u16 cdf[3];
        int idx = smulbb(9, shl(sigtype, 1) + QuantOffsetType) + RateLevelIndex;
        cdf[0] = 0; cdf[1] = SB_T(sign_cdf)[idx]; cdf[2] = 65535;
        for (int i = 0; i < FRAME; i++) if (q[i] != 0) rc_encode(rc, (q[i] >> 15) + 1, cdf);
    }
}

// SKP_Silk_encode_parameters (encode_parameters.c:33-182) for description `md` (0 or 1), 8 kHz.
SB_FN void encode_parameters(RangeEnc* rc, EncSilk* st, const EncCtrl* c, int md, int frame_in_packet, int vadFlag, const i8* q) {
    if (frame_in_packet == 0) {
        if (st->useMDIndex == 1) rc_encode(rc, md, SB_T(md_index_cdf));
        int i;
        for (i = 0; i < 3; i++) if (SB_T(sampling_rates_table)[i] == 8) break;
        rc_encode(rc, i, SB_T(sampling_rates_cdf));
    }
    int typeOffset = 2 * c->sigtype + c->QuantOffsetType;
    int Ix = st->typeOffsetPrev_md[md];
    if (frame_in_packet == 0) rc_encode(rc, typeOffset, SB_T(type_offset_cdf));
    else rc_encode(rc, typeOffset, &SB_T(type_offset_joint_cdf)[Ix * 5]);
    st->typeOffsetPrev_md[md] = typeOffset;
    if (frame_in_packet == 0) rc_encode(rc, c->GainsIndices[0], &SB_T(gain_cdf)[c->sigtype * 65]);
    else rc_encode(rc, c->GainsIndices[0], SB_T(delta_gain_cdf));
    for (int i = 1; i < NB_SUBFR; i++) rc_encode(rc, c->GainsIndices[i], SB_T(delta_gain_cdf));
    if (frame_in_packet == 0) rc_encode(rc, c->DeltaGainsIndices, SB_T(md_delta_gain_cdf));
    {
        const u16* cdf = c->sigtype == 0 ? SB_T(nlsf_cb0_cdf) : SB_T(nlsf_cb1_cdf);
        const i32* start = c->sigtype == 0 ? SB_T(nlsf_cb0_cdf_start) : SB_T(nlsf_cb1_cdf_start);
        for (int k = 0; k < 6; k++) rc_encode(rc, c->NLSFIndices[k], &cdf[start[k]]);
    }
    rc_encode(rc, c->NLSFInterpCoef_Q2, SB_T(nlsf_interp_cdf));
    if (c->sigtype == 0) {
        rc_encode(rc, c->lagIndex, SB_T(pitch_lag_nb_cdf));
        rc_encode(rc, c->contourIndex, SB_T(pitch_contour_nb_cdf));
        rc_encode(rc, c->PERIndex, SB_T(ltp_per_index_cdf));
        const u16* lc = c->PERIndex == 0 ? SB_T(ltp_cdf0) : (c->PERIndex == 1 ? SB_T(ltp_cdf1) : SB_T(ltp_cdf2));
        for (int k = 0; k < NB_SUBFR; k++) rc_encode(rc, c->LTPIndex[k], lc);
        rc_encode(rc, c->LTP_scaleIndex, SB_T(ltpscale_cdf));
    }
    rc_encode(rc, c->Seed, SB_T(seed_cdf));
    encode_pulses(rc, c->sigtype, c->QuantOffsetType, q);
    rc_encode(rc, vadFlag, SB_T(vadflag_cdf));
}

}  // namespace sb
