/*
 * File-to-file SOLO encoder / decoder on top of libsolo_b200.so, using only the reference's six-function API
 * (include/AGR_JC1_SDK_API.h) plus the framing helpers of include/solo_b200.h.  It plays the role of the reference's
 * test drivers (test/enc_main.c and test/dec_main.c of both source trees): same .bit file format, same loss simulation, so the files it
 * writes can be compared byte for byte with those of the reference CLI.
 *
 *   jc1_file_codec enc in.pcm out.bit [rate_bps] [dtx] [framesize_ms]
 *   jc1_file_codec dec in.bit out.pcm [loss_percent] [mode] [framesize_ms]     mode: 0 = both descriptions, 1 = MD1 only, 2 = MD2 only
 *
 * Build:  gcc -O2 -I include examples/jc1_file_codec.c -L solo_b200 -lsolo_b200 -Wl,-rpath,$PWD/solo_b200 -o jc1_file_codec
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "AGR_JC1_SDK_API.h"
#include "solo_b200.h"

#define MAX_ROW 1024

static int run_enc(int argc, char **argv) {
    USER_Ctrl_enc ec;
    memset(&ec, 0, sizeof ec);
    ec.targetRate_bps = argc > 4 ? atoi(argv[4]) : 13600;
    ec.samplerate = 16000;
    ec.dtx_enable = argc > 5 ? atoi(argv[5]) : 0;
    ec.framesize_ms = argc > 6 ? atoi(argv[6]) : 40;
    const int spp = 16 * ec.framesize_ms;
    FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
    if (!fi || !fo) { fprintf(stderr, "cannot open files\n"); return 2; }
    void *enc = AGR_Sate_Encoder_Init(&ec);
    if (!enc) { fprintf(stderr, "encoder init failed: %s\n", solo_b200_last_error()); return 3; }
    short pcm[640], nb[6];
    unsigned char row[MAX_ROW], rec[MAX_ROW + 4];
    long packets = 0, bytes = 0;
    while (fread(pcm, sizeof(short), (size_t)spp, fi) == (size_t)spp) {
        memset(nb, 0, sizeof nb);
        if (AGR_Sate_Encoder_Encode(enc, pcm, row, MAX_ROW, nb) < 0) { fprintf(stderr, "encode failed\n"); return 4; }
        int n = solo_b200_bitfile_pack(row, nb, rec, (int)sizeof rec);   /* int16 total, int16 len(MD2)+HB, payload */
        if (n < 0 || fwrite(rec, 1, (size_t)n, fo) != (size_t)n) { fprintf(stderr, "write failed\n"); return 5; }
        packets++; bytes += nb[0];
    }
    AGR_Sate_Encoder_Uninit(enc);
    fclose(fi); fclose(fo);
    fprintf(stderr, "%ld packets, %.1f bytes/packet\n", packets, packets ? (double)bytes / packets : 0.0);
    return 0;
}

/* the reference driver's loss process: one LCG draw per description and packet (dec_main.c:24,229-241) */
static unsigned lcg(unsigned s) { return 907633515u + s * 196314165u; }

static int run_dec(int argc, char **argv) {
    USER_Ctrl_dec dc;
    memset(&dc, 0, sizeof dc);
    dc.samplerate = 16000;
    const float loss = argc > 4 ? (float)atof(argv[4]) : 0.0f;
    const int mode = argc > 5 ? atoi(argv[5]) : 0;
    dc.framesize_ms = argc > 6 ? atoi(argv[6]) : 40;
    FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
    if (!fi || !fo) { fprintf(stderr, "cannot open files\n"); return 2; }
    void *dec = AGR_Sate_Decoder_Init(&dc);
    if (!dec) { fprintf(stderr, "decoder init failed: %s\n", solo_b200_last_error()); return 3; }
    unsigned char hdr[4], row[MAX_ROW], arg[MAX_ROW];
    short pcm[960], ns = 0;
    unsigned seed = 1;
    long packets = 0;
    while (fread(hdr, 1, 4, fi) == 4) {
        short nb[2] = {(short)(hdr[0] | (hdr[1] << 8)), (short)(hdr[2] | (hdr[3] << 8))};
        if (nb[0] < 0 || nb[0] > MAX_ROW || fread(row, 1, (size_t)nb[0], fi) != (size_t)nb[0]) break;
        /* sender side: two network packets; receiver side: whatever survived the channel */
        const unsigned char *p1, *p2;
        int n1, n2, lost[2];
        if (solo_b200_split_packet(row, nb, &p1, &n1, &p2, &n2)) { fprintf(stderr, "bad record\n"); return 4; }
        for (int j = 0; j < 2; j++) {
            seed = lcg(seed);
            float v = (float)(((int)seed >> 16) + (1 << 15)) / 65535.0f;
            lost[j] = !(v >= loss / 100.0f) || (j == 0 ? n1 : n2) == 0;
        }
        if (mode == 1) { lost[0] = n1 == 0; lost[1] = 1; }
        if (mode == 2) { lost[0] = 1; lost[1] = n2 == 0; }
        short anb[6] = {0, 0, 0, 0, 0, 0};
        int flag = 0;
        if (solo_b200_merge_packets(lost[0] ? NULL : p1, lost[0] ? 0 : n1, lost[1] ? NULL : p2, lost[1] ? 0 : n2, arg, MAX_ROW, anb, &flag)) return 5;
        int r = AGR_Sate_Decoder_Decode(dec, pcm, &ns, arg, anb, flag);
        if (r < 0) { fprintf(stderr, "decode failed (%d) at packet %ld\n", r, packets); return 6; }
        fwrite(pcm, sizeof(short), (size_t)ns, fo);
        packets++;
    }
    AGR_Sate_Decoder_Uninit(dec);
    fclose(fi); fclose(fo);
    fprintf(stderr, "%ld packets decoded\n", packets);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "enc")) return run_enc(argc, argv);
    if (argc >= 4 && !strcmp(argv[1], "dec")) return run_dec(argc, argv);
    fprintf(stderr, "usage: %s enc in.pcm out.bit [rate_bps] [dtx] [framesize_ms]\n       %s dec in.bit out.pcm [loss_percent] [mode] [framesize_ms]\n", argv[0], argv[0]);
    return 1;
}
