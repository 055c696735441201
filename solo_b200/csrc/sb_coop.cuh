// solo_b200 -- encoder analysis stage in its warp-cooperative form: ONE stream = ONE warp.
//
// Every routine here is called by all 32 lanes of the warp that owns the stream (convergently), works on arrays in shared
// memory and keeps scalars in registers.  The work of a routine is laid out over the lanes in one of four ways:
//   * lanes over outputs      -- FIR filters, correlations per lag, code-vector errors: no communication at all;
//   * lanes over terms + sum  -- wrap-around integer sums (mod 2^32 / 2^64 addition is associative, so the bits match the
//                                reference's sequential accumulation);
//   * lanes over stages       -- the warped all-pass chains run as a skewed wavefront (lane = stage, one shuffle per step);
//   * lane per instance       -- short recursions that exist several times per frame (4 shaping windows, 4 LTP sub-frames,
//                                4 interpolation candidates) run as one scalar instance per lane.
// What is genuinely one serial recurrence at signal rate (VAD filter bank, high-pass biquad, the 2:1 decimator, the
// low-frequency shaping recursion) runs on lane 0.
// The scalar routines of sb_enc_*.cuh / sb_sigproc.cuh remain the executable specification: tests/hostsim compiles both and
// checks this file (32 fibres per stream, sb_par.cuh SB_EMU) against the golden bitstream on a machine without a GPU.
#pragma once
#include "sb_enc_front.cuh"
#include "sb_enc_pred.cuh"
#include "sb_enc_shape.cuh"
#include "sb_par.cuh"

#if SB_COOP_ACTIVE
namespace sb {

// ---------------------------------------------------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------------------------------------------------
SB_HD int popc32(u32 x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
SB_HD int ctz32(u32 x) {   // x != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// Zero-state FIR (SKP_Silk_MA_Prediction / SKP_Silk_LPC_analysis_filter with a cleared state, SKP_Silk_MA.c:41-118):
// out[k] = sat16(rshift_round((in[k] << 12) - sum_{d < min(ORD, k)} B[d] * in[k-1-d], 12)); lanes over outputs.
template <int ORD, bool SAT> SB_FN void c_fir_zero_state(const i16* in, const i16* B_Q12, i16* out, int len) {
    i32 b[ORD];
#pragma unroll
    for (int d = 0; d < ORD; d++) b[d] = B_Q12[d];
    SB_PARFOR(k, 0, len) {
        i32 acc = 0;
        if (k >= ORD) {
#pragma unroll
            for (int d = 0; d < ORD; d++) acc = addw(acc, (i32)in[k - 1 - d] * b[d]);
        } else {
#pragma unroll
            for (int d = 0; d < ORD; d++) if (d < k) acc = addw(acc, (i32)in[k - 1 - d] * b[d]);
        }
        const i32 x = in[k];
        const i32 o = SAT ? sub_sat32(shl(x, 12), acc) : subw(shl(x, 12), acc);
        out[k] = (i16)sat16(rshift_round(o, 12));
    }
}

// SKP_Silk_sum_sqr_shift (sum_sqr_shift.c:40-98) over the whole warp.  The reference accumulates pair by pair and starts
// shifting once the running sum reaches bit 31; the running sums are monotone, so when the exact total stays below 2^31
// no shift ever happened and the result follows from the total (lanes over terms).  Otherwise lane 0 replays the scalar
// routine.  All lanes return the same values.
SB_FN void c_sum_sqr_shift(i32* energy, i32* shift, const i16* x, int len, int odd_start) {
    i64 part = 0;
    SB_PARFOR(i, 0, len) part += (i64)((i32)x[i] * (i32)x[i]);
    const i64 total = wsum64(part);
    if (total < ((i64)1 << 31)) {      // uniform branch
        i32 nrg = (i32)total, shft = 0;
        if (nrg & 0xC0000000) { nrg = (i32)((u32)nrg >> 2); shft = 2; }
        *energy = nrg; *shift = shft;
        return;
    }
    i32 e = 0, s = 0;
    if (SB_LANE0) sum_sqr_shift(&e, &s, x, len, odd_start);
    *energy = wbcast(e, 0);
    *shift = wbcast(s, 0);
}

// SKP_Silk_schur (schur.c:40-93), order <= 31: lane n keeps C[n][1] and C[n+k+1][0] in registers; the second slides down
// one lane per step.  c: correlations (shared memory); rc_Q15: shared; returns the residual energy on every lane.
SB_FN i32 c_schur(i16* rc_Q15, const i32* c, int order) {
    const int lane = SB_LANE;
    const int lz = clz32(c[0]);
    i32 a = 0, b = 0;
    if (lane <= order) { const i32 v = c[lane]; a = lz < 2 ? (v >> 1) : (lz > 2 ? shl(v, lz - 2) : v); }
    b = wshfl_down(a, 1);
    for (int k = 0; k < order; k++) {
        const i32 c00 = wbcast(b, 0), c01 = wbcast(a, 0);
        const i32 rc = sat16(negw(c00 / imax(c01 >> 15, 1)));
        if (lane == 0) rc_Q15[k] = (i16)rc;
        if (lane < order - k) {
            const i32 t1 = b, t2 = a;
            b = smlawb(t1, shl(t2, 1), rc);
            a = smlawb(t2, shl(t1, 1), rc);
        }
        b = wshfl_down(b, 1);
    }
    return wbcast(a, 0);
}
// SKP_Silk_schur64 (schur64.c:42-91)
SB_FN i32 c_schur64(i32* rc_Q16, const i32* c, int order) {
    const int lane = SB_LANE;
    if (c[0] <= 0) {        // uniform
        if (lane < order) rc_Q16[lane] = 0;
        return 0;
    }
    i32 a = lane <= order ? c[lane] : 0;
    i32 b = wshfl_down(a, 1);
    for (int k = 0; k < order; k++) {
        const i32 c00 = wbcast(b, 0), c01 = wbcast(a, 0);
        const i32 rc_Q31 = div32_varq(negw(c00), c01, 31);
        if (lane == 0) rc_Q16[k] = rshift_round(rc_Q31, 15);
        if (lane < order - k) {
            const i32 t1 = b, t2 = a;
            b = addw(t1, smmul(shl(t2, 1), rc_Q31));
            a = addw(t2, smmul(shl(t1, 1), rc_Q31));
        }
        b = wshfl_down(b, 1);
    }
    return wbcast(a, 0);
}
// SKP_Silk_k2a (k2a.c:40-60): lane n keeps A[n]; returns it (valid for n < order).  rc_Q15: shared memory.
SB_FN i32 c_k2a(const i16* rc_Q15, int order) {
    const int lane = SB_LANE;
    i32 A = 0;
    for (int k = 0; k < order; k++) {
        const i32 rc = rc_Q15[k];
        const i32 t = wshfl(A, (k - lane - 1) & 31);
        if (lane < k) A = smlawb(A, shl(t, 1), rc);
        if (lane == k) A = negw(shl(rc, 9));
    }
    return A;
}
// SKP_Silk_k2a_Q16 (k2a_Q16.c:40-60)
SB_FN i32 c_k2a_q16(const i32* rc_Q16, int order) {
    const int lane = SB_LANE;
    i32 A = 0;
    for (int k = 0; k < order; k++) {
        const i32 rc = rc_Q16[k];
        const i32 t = wshfl(A, (k - lane - 1) & 31);
        if (lane < k) A = smlaww(A, t, rc);
        if (lane == k) A = negw(shl(rc, 8));
    }
    return A;
}
// chirp factor that SKP_Silk_bwexpander (bwexpander.c:31-48) applies to coefficient `lane` (d coefficients)
SB_FN i32 c_bwexpander_chirp16(int d, i32 chirp_Q16) {
    const int lane = SB_LANE;
    const i32 cm1 = chirp_Q16 - 65536;
    i32 mine = chirp_Q16;
    for (int i = 0; i < d - 1; i++) {
        if (lane == i) mine = chirp_Q16;
        chirp_Q16 += rshift_round(mulw(chirp_Q16, cm1), 16);
    }
    if (lane >= d - 1) mine = chirp_Q16;
    return mine;
}
// factor that SKP_Silk_bwexpander_32 (bwexpander_32.c:31-47) multiplies into coefficient `lane`
SB_FN i32 c_bwexpander32_factor(int d, i32 chirp_Q16) {
    const int lane = SB_LANE;
    i32 t = chirp_Q16, mine = chirp_Q16;
    for (int i = 0; i < d - 1; i++) {
        if (lane == i) mine = t;
        t = smulww(chirp_Q16, t);
    }
    if (lane >= d - 1) mine = t;
    return mine;
}

// ---------------------------------------------------------------------------------------------------------------------
// pitch analysis (SKP_Silk_find_pitch_lags_FIX.c:32-125, SKP_Silk_pitch_analysis_core.c:65-560 at 8 kHz, complexity 2)
// ---------------------------------------------------------------------------------------------------------------------
struct PitchScr {
    i16 Wsig[PITCH_LPC_WIN];
    i16 sig8[2 * FRAME];
    i16 sig4[FRAME];
    i16 C1[2][66];        // first-stage correlations, index lag - 8
    i16 C2[4][48];        // second-stage correlations by candidate slot
    i16 d_comp[48];
    i16 dc[160];          // candidate marks / their running sums, index lag
    i32 d_srch[24];
    u8 slot[152];         // lag -> candidate slot + 1 (0: not computed, correlation 0)
    i32 acorr[12];
    i16 rc_Q15[16];
    i16 A_Q12[16];
};

SB_FN int c_pitch_analysis_core(PitchScr* P, const i16* signal, i32* pitch_out, i32* lagIndex, i32* contourIndex, i32* LTPCorr_Q15,
                                i32 prevLag, i32 search_thres1_Q16, i32 search_thres2_Q15) {
    enum { FL8 = 320, FL4 = 160, SF8 = 40, MINL8 = 16, MAXL8 = 144, MINL4 = 8, MAXL4 = 72, NCB = 11, NL4 = MAXL4 - MINL4 + 1 };
    const int lane = SB_LANE;
    i16* sig8 = P->sig8;
    i16* sig4 = P->sig4;
    SB_PARFOR(i, 0, FL8) sig8[i] = signal[i];
    SB_SYNC();
    // 2:1 decimator: one recurrence over the frame (lane 0)
    if (lane == 0) { i32 fs[2] = {0, 0}; resampler_down2(fs, sig4, sig8, FL8); }
    SB_SYNC();
    // low-pass (descending in-place loop of the reference = old values on the right-hand side) + scaling
    i32 v[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { const int i = lane + 32 * j; v[j] = i > 0 ? add_sat16(sig4[i], sig4[i - 1]) : (i32)sig4[0]; }
    SB_SYNC();
    i32 mx = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) mx = imax(mx, v[j] * v[j]);
    mx = wmax(mx);
    i32 shift;
    // pitch_find_scaling(signal_4kHz, 160, 80): the reference takes |x| at the largest x^2 (-32768 counts as 32767)
    const i32 x_max4 = mx >= 1073676289 ? 32767 : wmax(imax(imax(imax(iabs(v[0]), iabs(v[1])), imax(iabs(v[2]), iabs(v[3]))), iabs(v[4])));
    {
        i32 nbits = x_max4 < 32767 ? 32 - clz32(smulbb(x_max4, x_max4)) : 30;
        nbits += 17 - (clz32(imax(SF8, FL4 >> 1)) - 16);
        shift = nbits < 31 ? 0 : nbits - 30;
    }
#pragma unroll
    for (int j = 0; j < 5; j++) sig4[lane + 32 * j] = (i16)((i16)v[j] >> shift);
    SB_SYNC();

    // ---- first stage at 4 kHz: lanes over (half, lag) ----
    const i16* target0 = &sig4[FL4 >> 1];
    i32 e8 = 0;     // energy of the first basis vector (lag 8) of half `lane` (lanes 0, 1)
    if (lane < 2) { const i16* b = target0 + lane * SF8 - MINL4; for (int i = 0; i < SF8; i++) e8 = addw(e8, (i32)b[i] * (i32)b[i]); }
    const i32 e8_0 = wbcast(e8, 0), e8_1 = wbcast(e8, 1);
    SB_PARFOR(t, 0, 2 * NL4) {
        const int k = t >= NL4, d = MINL4 + t - k * NL4;
        const i16* target = target0 + k * SF8;
        const i16* basis = target - d;
        i32 cross = 0, e = 0;
#pragma unroll 8
        for (int i = 0; i < SF8; i++) { const i32 bv = basis[i]; cross = addw(cross, (i32)target[i] * bv); e = addw(e, bv * bv); }
        const i32 ek = k ? e8_1 : e8_0;
        const i32 normalizer = addw(add_sat32(ek, smulbb(SF8, 4000)), subw(e, ek));   // recursion of the reference, unrolled
        P->C1[k][d - MINL4] = (i16)sat16(cross / (sqrt_approx(normalizer) + 1));
    }
    SB_SYNC();
    i32 cs[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int i = lane + 32 * j;
        if (i < NL4) {
            i32 sum = ((i32)P->C1[0][i] + (i32)P->C1[1][i]) >> 1;
            cs[j] = smlawb(sum, sum, shl(-(i + MINL4), 4));
        }
    }
    SB_SYNC();
#pragma unroll
    for (int j = 0; j < 3; j++) { const int i = lane + 32 * j; if (i < NL4) P->C1[0][i] = (i16)cs[j]; }
    SB_SYNC();
    int length_d_srch = 4 + 2 * 2;
    if (lane == 0) insertion_sort_decreasing_i16(&P->C1[0][0], P->d_srch, NL4, length_d_srch);
    i32 energy;
    {
        i32 part = 0;
        SB_PARFOR(i, 0, FL4 >> 1) part = addw(part, (i32)target0[i] * (i32)target0[i]);
        energy = add_pos_sat32(wsum(part), 1000);
    }
    SB_SYNC();
    const i32 Cmax = P->C1[0][0];
    i32 threshold = smulbb(Cmax, Cmax);
    if ((energy >> (4 + 2)) > threshold) {     // uniform
        if (lane < 4) pitch_out[lane] = 0;
        if (lane == 0) { *LTPCorr_Q15 = 0; *lagIndex = 0; *contourIndex = 0; }
        SB_SYNC();
        return 1;
    }
    threshold = smulwb(search_thres1_Q16, Cmax);
    {
        const bool ok = lane < length_d_srch && P->C1[0][lane] > threshold;
        const u32 bad = ~wballot(ok) & ((1u << length_d_srch) - 1);
        if (bad) length_d_srch = ctz32(bad);         // the list ends at the first entry below the threshold
    }
    SB_PARFOR(i, 0, 160) P->dc[i] = 0;
    SB_SYNC();
    if (lane < length_d_srch) { const i32 d = (P->d_srch[lane] + MINL4) << 1; P->dc[d] = 1; }
    SB_SYNC();
    // running sums over 3, then over 4 lags (the reference's in-place descending loops read unmodified lower entries)
    i32 w5[5];
#pragma unroll
    for (int j = 0; j < 5; j++) { const int i = MINL8 + lane + 32 * j; w5[j] = i <= MAXL8 + 3 ? P->dc[i] + P->dc[i - 1] + P->dc[i - 2] : 0; }
    SB_SYNC();
#pragma unroll
    for (int j = 0; j < 5; j++) { const int i = MINL8 + lane + 32 * j; if (i <= MAXL8 + 3) P->dc[i] = (i16)w5[j]; }
    SB_SYNC();
    length_d_srch = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int i = MINL8 + lane + 32 * j;
        const bool on = i < MAXL8 + 1 && P->dc[i + 1] > 0;
        const u32 m = wballot(on);
        if (on) P->d_srch[length_d_srch + popc32(m & ((1u << lane) - 1))] = i;
        length_d_srch += popc32(m);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) { const int i = MINL8 + lane + 32 * j; w5[j] = i <= MAXL8 + 3 ? P->dc[i] + P->dc[i - 1] + P->dc[i - 2] + P->dc[i - 3] : 0; }
    SB_SYNC();
    int length_d_comp = 0;
    SB_PARFOR(i, 0, 152) P->slot[i] = 0;
    SB_SYNC();
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int i = MINL8 + lane + 32 * j;
        const bool on = i < MAXL8 + 4 && w5[j] > 0;
        const u32 m = wballot(on);
        if (on) {
            const int q = length_d_comp + popc32(m & ((1u << lane) - 1));
            P->d_comp[q] = (i16)(i - 2);
            P->slot[i - 2] = (u8)(q + 1);
        }
        length_d_comp += popc32(m);
    }

    // ---- second stage at 8 kHz: lanes over (sub-frame, candidate lag) ----
    {
        i32 m8 = 0;
        SB_PARFOR(i, 0, FL8) m8 = imax(m8, iabs((i32)sig8[i]));
        const i32 x_max = imin(wmax(m8), 32767);
        i32 nbits = x_max < 32767 ? 32 - clz32(smulbb(x_max, x_max)) : 30;
        nbits += 17 - (clz32(SF8) - 16);
        shift = nbits < 31 ? 0 : nbits - 30;
    }
    if (shift > 0) SB_PARFOR(i, 0, FL8) sig8[i] = (i16)(sig8[i] >> shift);
    SB_SYNC();
    i32 et = 0;      // energy of target sub-frame `lane` (lanes 0..3)
    if (lane < 4) { const i16* tp = &sig8[FL4 + lane * SF8]; for (int i = 0; i < SF8; i++) et = addw(et, (i32)tp[i] * (i32)tp[i]); }
    const i32 et0 = wbcast(et, 0), et1 = wbcast(et, 1), et2 = wbcast(et, 2), et3 = wbcast(et, 3);
    SB_PARFOR(t, 0, 4 * length_d_comp) {
        const int k = t / length_d_comp, j = t - k * length_d_comp;
        const int d = P->d_comp[j];
        const i16* target = &sig8[FL4 + k * SF8];
        const i16* basis = target - d;
        i32 cross_corr = 0, energy_basis = 0;
#pragma unroll 8
        for (int i = 0; i < SF8; i++) { const i32 bv = basis[i]; cross_corr = addw(cross_corr, (i32)target[i] * bv); energy_basis = addw(energy_basis, bv * bv); }
        const i32 energy_target = k == 0 ? et0 : (k == 1 ? et1 : (k == 2 ? et2 : et3));
        i32 cv = 0;
        if (cross_corr > 0) {
            i32 en = imax(energy_target, energy_basis);
            int lz = clz32(cross_corr);
            int lshift = limit(lz - 1, 0, 15);
            i32 temp32 = shl(cross_corr, lshift) / ((en >> (15 - lshift)) + 1);
            temp32 = smulwb(cross_corr, temp32);
            temp32 = add_sat32(temp32, temp32);
            lz = clz32(temp32);
            lshift = limit(lz - 1, 0, 15);
            en = imin(energy_target, energy_basis);
            cv = (i16)(shl(temp32, lshift) / ((en >> (15 - lshift)) + 1));
        }
        P->C2[k][j] = (i16)cv;
    }
    SB_SYNC();
    // ---- search over the short list (one candidate lag per lane) ----
    const i16* cbl = SB_T(pitch_cb_lags_stage2);   // [4][11]
    const i32 prevLag_log2_Q7 = prevLag > 0 ? lin2log(prevLag) : 0;
    const i32 corr_thres_Q15 = smulbb(search_thres2_Q15, search_thres2_Q15) >> 13;
    const i32 ltpcorr_prev = *LTPCorr_Q15;
    i32 key = SB_I32_MIN, my_cc = SB_I32_MIN, my_cb = 0, my_d = 0;
    if (lane < length_d_srch) {
        const int d = P->d_srch[lane];
        i32 CCmax_new = SB_I32_MIN; int CBimax_new = 0;
        for (int j = 0; j < NCB; j++) {
            i32 cc = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { const int sl = P->slot[d + cbl[i * NCB + j]]; cc += sl ? (i32)P->C2[i][sl - 1] : 0; }
            if (cc > CCmax_new) { CCmax_new = cc; CBimax_new = j; }
        }
        const i32 lag_log2_Q7 = lin2log(d);
        i32 CCmax_new_b = CCmax_new - (smulbb(4 * 6554, lag_log2_Q7) >> 7);
        if (prevLag > 0) {
            i32 dl = lag_log2_Q7 - prevLag_log2_Q7;
            dl = smulbb(dl, dl) >> 7;
            i32 prev_lag_bias_Q15 = smulbb(4 * 6554, ltpcorr_prev) >> 15;
            prev_lag_bias_Q15 = mulw(prev_lag_bias_Q15, dl) / (dl + (1 << 6));
            CCmax_new_b -= prev_lag_bias_Q15;
        }
        if (CCmax_new > corr_thres_Q15 && cbl[0 * NCB + CBimax_new] <= MINL8) key = CCmax_new_b;
        my_cc = CCmax_new; my_cb = CBimax_new; my_d = d;
    }
    i32 best = key, who = lane;
    wargmax(best, who);        // first of equal maxima, as the ascending scan with a strict comparison keeps it
    if (best == SB_I32_MIN) {  // uniform
        if (lane < 4) pitch_out[lane] = 0;
        if (lane == 0) { *LTPCorr_Q15 = 0; *lagIndex = 0; *contourIndex = 0; }
        SB_SYNC();
        return 1;
    }
    const i32 CCmax = imax(wshfl(my_cc, who), 0);
    const i32 CBimax = wshfl(my_cb, who), lag = wshfl(my_d, who);
    if (lane < 4) pitch_out[lane] = lag + cbl[lane * NCB + CBimax];
    if (lane == 0) { *LTPCorr_Q15 = sqrt_approx(shl(CCmax, 13)); *lagIndex = lag - MINL8; *contourIndex = CBimax; }
    SB_SYNC();
    return 0;
}

// x points at x_buf + FRAME; res receives 336 samples of LPC residual.  All lanes return (and c gets) the signal type.
SB_FN void c_find_pitch_lags(EncSilk* st, EncCtrl* c, PitchScr* P, i16* res, const i16* x) {
    enum { BUF_LEN = LA_PITCH + 2 * FRAME, ORD = 10 };
    const int lane = SB_LANE;
    const i16* x_buf = x - FRAME;
    const i16* x_buf_ptr = x_buf + BUF_LEN - PITCH_LPC_WIN;
    // 24 ms window: sine slopes on two lanes (16 samples each), flat part copied by all
    if (lane == 0) apply_sine_window(P->Wsig, x_buf_ptr, 1, LA_PITCH);
    if (lane == 1) apply_sine_window(P->Wsig + PITCH_LPC_WIN - LA_PITCH, x_buf_ptr + PITCH_LPC_WIN - LA_PITCH, 2, LA_PITCH);
    SB_PARFOR(i, 0, PITCH_LPC_WIN - 2 * LA_PITCH) P->Wsig[LA_PITCH + i] = x_buf_ptr[LA_PITCH + i];
    SB_SYNC();
    // autocorrelation, 11 lags (SKP_Silk_autocorr, autocorr.c:40-77): lanes over samples, 64-bit partial sums
    i64 acc[ORD + 1];
#pragma unroll
    for (int i = 0; i <= ORD; i++) acc[i] = 0;
    for (int j = 0; j < PITCH_LPC_WIN / 32; j++) {
        const int n = lane + 32 * j;
        const i32 xn = P->Wsig[n];
#pragma unroll
        for (int i = 0; i <= ORD; i++) if (n + i < PITCH_LPC_WIN) acc[i] += (i64)(xn * (i32)P->Wsig[n + i]);
    }
#pragma unroll
    for (int i = 0; i <= ORD; i++) acc[i] = wsum64(acc[i]);
    {
        const i64 corr64 = acc[0] + 1;
        const int nrs = 35 - clz64(corr64);
#pragma unroll
        for (int i = 0; i <= ORD; i++) {
            const i64 s = i == 0 ? corr64 : acc[i];
            const i32 r = nrs <= 0 ? shl((i32)s, -nrs) : (i32)(s >> nrs);
            if (lane == i) P->acorr[i] = i == 0 ? smlawb(r, r, SB_FIXC(1e-3f, 16)) : r;
        }
    }
    SB_SYNC();
    const i32 ac0 = P->acorr[0];
    const i32 res_nrg = c_schur(P->rc_Q15, P->acorr, ORD);
    const i32 predGain = div32_varq(ac0, imax(res_nrg, 1), 16);
    SB_SYNC();
    const i32 A_Q24 = c_k2a(P->rc_Q15, ORD);
    {
        const i32 chirp = c_bwexpander_chirp16(ORD, SB_FIXC(0.99f, 16));
        if (lane < ORD) P->A_Q12[lane] = (i16)rshift_round(mulw(chirp, (i32)(i16)sat16(A_Q24 >> 12)), 16);
    }
    SB_SYNC();
    c_fir_zero_state<ORD, false>(x_buf, P->A_Q12, res, BUF_LEN);
    SB_SYNC();
    if (lane < ORD) res[lane] = 0;
    i32 thrhld_Q15 = SB_FIXC(0.45, 15);
    thrhld_Q15 = smlabb(thrhld_Q15, SB_FIXC(-0.004, 15), ORD);
    thrhld_Q15 = smlabb(thrhld_Q15, SB_FIXC(-0.1, 7), st->speech_activity_Q8);
    thrhld_Q15 = smlabb(thrhld_Q15, SB_FIXC(0.15, 15), st->prev_sigtype);
    thrhld_Q15 = smlawb(thrhld_Q15, SB_FIXC(-0.1, 16), c->input_tilt_Q15);
    thrhld_Q15 = sat16(thrhld_Q15);
    const i32 prevLag = st->prevLag;
    SB_SYNC();
    const int sigtype = c_pitch_analysis_core(P, res, c->pitchL, &c->lagIndex, &c->contourIndex, &st->LTPCorr_Q15, prevLag,
                                              SB_FIXC(0.7f, 16), (i16)thrhld_Q15);
    if (lane == 0) { c->sigtype = sigtype; c->predGain_Q16 = predGain; }
    SB_SYNC();
}

// ---------------------------------------------------------------------------------------------------------------------
// frame / packet drivers
// ---------------------------------------------------------------------------------------------------------------------
// Working set of one stream during a packet (shared memory in the warp-per-stream kernel).
struct CoopWork {
    alignas(16) i16 low[2 * FRAME];        // low band of the packet (both 20 ms frames)
    i16 pIn_HP[FRAME];
    i16 res_pitch[2 * FRAME + LA_PITCH];
    EncCtrl c;
    alignas(16) i16 xfw[FRAME];
    i32 vadFlag;
    union {
        PitchScr pitch;
        i16 vadX[4 * (FRAME / 2)];
    } u;
};

// SKP_Silk_encode_frame_FIX (encode_frame_FIX.c:34-131, 151-165, 199-208) up to the quantiser, one 20 ms frame.
SB_FN void c_encode_frame_analysis(EncSilk* st, CoopWork* W, const i16* pIn, int frame_in_packet) {
    EncCtrl* c = &W->c;
    i16* x_frame = st->x_buf + FRAME;
    SB_SERIAL(
        c->Seed = st->frameCounter++ & 3;
        vad_get_sa_q8(&st->vad, &st->speech_activity_Q8, c->input_quality_bands_Q15, &c->input_tilt_Q15, pIn);
        hp_variable_cutoff(st, c, W->pIn_HP, pIn);
    );
    SB_PARFOR(i, 0, FRAME) x_frame[LA_SHAPE + i] = W->pIn_HP[i];   // LP_variable_cutoff is a copy (transition_frame_no == 0)
    SB_SYNC();
    c_find_pitch_lags(st, c, &W->u.pitch, W->res_pitch, x_frame);
    SB_SERIAL(
        noise_shape_analysis(st, c, W->res_pitch + FRAME, x_frame);
        prefilter(st, c, W->xfw, x_frame);
        find_pred_coefs(st, c, W->res_pitch, frame_in_packet, nullptr);
        process_gains(st, c, frame_in_packet);
        if (st->speech_activity_Q8 < SB_FIXC(0.1f, 8)) {
            st->vadFlag = 0;
            st->noSpeechCounter++;
            if (st->noSpeechCounter > 5) st->inDTX = 1;
            if (st->noSpeechCounter > 20 + 5) { st->noSpeechCounter = 5; st->inDTX = 0; }
        } else {
            st->noSpeechCounter = 0; st->inDTX = 0; st->vadFlag = 1;
        }
        W->vadFlag = st->vadFlag;
        st->prev_sigtype = c->sigtype;
        st->prevLag = c->pitchL[NB_SUBFR - 1];
        st->first_frame_after_reset = 0;
    );
    {   // x_buf slides by one frame (old values on the right-hand side)
        i32 keep[(FRAME + LA_SHAPE) / 2 / 32 + 1];
        const i32* src = reinterpret_cast<const i32*>(st->x_buf + FRAME);
        i32* dst = reinterpret_cast<i32*>(st->x_buf);
        int q = 0;
        SB_PARFOR(i, 0, (FRAME + LA_SHAPE) / 2) keep[q++] = src[i];
        SB_SYNC();
        q = 0;
        SB_PARFOR(i, 0, (FRAME + LA_SHAPE) / 2) dst[i] = keep[q++];
        SB_SYNC();
    }
}

// Stage A for the SILK core of one packet.  st, W: shared memory; W->low already holds the low band; scr: global memory.
SB_FN void c_enc_packet_analysis(EncSilk* st, CoopWork* W, EncScratch* scr) {
    const int nf = st->frames_per_packet;
    for (int f = 0; f < nf; f++) {
        c_encode_frame_analysis(st, W, W->low + f * FRAME, f);
        const i32* src = reinterpret_cast<const i32*>(&W->c);
        i32* dst = reinterpret_cast<i32*>(&scr->c[f]);
        SB_PARFOR(i, 0, (int)(sizeof(EncCtrl) / 4)) dst[i] = src[i];
        const i32* xs = reinterpret_cast<const i32*>(W->xfw);
        i32* xd = reinterpret_cast<i32*>(scr->xfw[f]);
        SB_PARFOR(i, 0, FRAME / 2) xd[i] = xs[i];
        if (SB_LANE0) scr->vadFlag[f] = W->vadFlag;
        SB_SYNC();
    }
    if (SB_LANE0) scr->dtx_drop = (st->useDTX && st->inDTX) ? 1 : 0;
}

}  // namespace sb
#endif  // SB_COOP_ACTIVE
