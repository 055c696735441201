/*
 * CPU baseline runner for the reference SOLO codec (test / measurement infrastructure).
 *
 * dlopen()s oracle/_ref/libjc1_fix.so (encoder, fixed-point build) and
 * oracle/_ref/libjc1_flp.so (decoder, float build) -- the UNMODIFIED reference compiled by
 * oracle/Makefile -- and drives one AGR_Sate_Encoder / AGR_Sate_Decoder handle pair per stream
 * through the public six-function API (reference: JC1_SDK_SRC_ARM/interface/AGR_JC1_SDK_API.h:33-64,
 * call pattern of JC1_SDK_SRC_ARM/test/enc_main.c:178-184 and JC1_SDK_SRC_FLP/test/dec_main.c:343).
 *
 * usage: cpu_baseline <dir-with-libs> <pcm-file> <threads> <streams-per-thread> <packets-per-stream> [rate_bps] [pin]
 * pin = 1: thread t is bound to the t-th CPU of the process's affinity mask (one thread per usable core).
 * Input is the "speech-replay" batch of SURVEY.md 8(d): stream s reads the clip circularly from
 * sample offset (s*7919*640) mod nsamples with gain 2^-(s mod 4).
 * Timing: wall clock over all threads, Init and the first packet of every stream excluded.
 * Prints one JSON line.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int mode, targetRate_bps, samplerate, dtx_enable, framesize_ms, joint_enable, joint_mode, useMDIndex; } enc_ctrl_t;
typedef struct { int packetLoss_perc, samplerate, framesize_ms, joint_enable, joint_mode, useMDIndex; } dec_ctrl_t;

typedef void *(*enc_init_f)(enc_ctrl_t *);
typedef int (*enc_encode_f)(void *, const short *, unsigned char *, int, short *);
typedef int (*enc_uninit_f)(void *);
typedef void *(*dec_init_f)(dec_ctrl_t *);
typedef int (*dec_decode_f)(void *, short *, short *, const unsigned char *, short *, int);
typedef int (*dec_uninit_f)(void *);

static enc_init_f enc_init; static enc_encode_f enc_encode; static enc_uninit_f enc_uninit;
static dec_init_f dec_init; static dec_decode_f dec_decode; static dec_uninit_f dec_uninit;

static short *clip; static long clip_n;
static int n_threads, streams_per_thread, packets, rate_bps;
static pthread_barrier_t bar;
static double t_start[256], t_end[256];
static unsigned long long bytes_out[256];
static unsigned long long pcm_hash[256];

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void fill(short *dst, long stream, int pkt) {
    long off = (long)(((long long)stream * 7919LL * 640LL) % clip_n) + (long)pkt * 640;
    int sh = (int)(stream & 3);
    for (int i = 0; i < 640; i++) dst[i] = (short)(clip[(off + i) % clip_n] >> sh);
}

static int pin_threads, n_allowed, allowed[1024];

static void *worker(void *arg) {
    long tid = (long)arg;
    if (pin_threads && n_allowed > 0) {
        cpu_set_t one; CPU_ZERO(&one); CPU_SET(allowed[tid % n_allowed], &one);
        pthread_setaffinity_np(pthread_self(), sizeof one, &one);
    }
    void **enc = malloc(sizeof(void *) * streams_per_thread), **dec = malloc(sizeof(void *) * streams_per_thread);
    for (int s = 0; s < streams_per_thread; s++) {
        enc_ctrl_t ec = {2, rate_bps, 16000, 0, 40, 0, 0, 0};
        dec_ctrl_t dc = {0, 16000, 40, 0, 0, 0};
        enc[s] = enc_init(&ec); dec[s] = dec_init(&dc);
    }
    short pcm[640], out[960], nb[6], nsamp; unsigned char bits[1024];
    unsigned long long nbytes = 0, h = 1469598103934665603ULL;
    for (int p = 0; p < packets + 1; p++) {
        if (p == 1) { pthread_barrier_wait(&bar); t_start[tid] = now(); }
        for (int s = 0; s < streams_per_thread; s++) {
            long sid = tid * streams_per_thread + s;
            fill(pcm, sid, p);
            memset(nb, 0, sizeof nb);
            int n = enc_encode(enc[s], pcm, bits, 1024, nb);
            short nbd[2] = {nb[0], nb[1]};
            dec_decode(dec[s], out, &nsamp, bits, nbd, 4);
            if (p) { nbytes += (unsigned)n; for (int i = 0; i < 640; i += 37) h = (h ^ (unsigned short)out[i]) * 1099511628211ULL; }
        }
    }
    t_end[tid] = now();
    bytes_out[tid] = nbytes; pcm_hash[tid] = h;
    for (int s = 0; s < streams_per_thread; s++) { enc_uninit(enc[s]); dec_uninit(dec[s]); }
    free(enc); free(dec);
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s libdir pcm threads streams_per_thread packets [rate]\n", argv[0]); return 2; }
    char path[1024];
    snprintf(path, sizeof path, "%s/libjc1_fix.so", argv[1]);
    void *hf = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    snprintf(path, sizeof path, "%s/libjc1_flp.so", argv[1]);
    void *hl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!hf || !hl) { fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 1; }
    enc_init = (enc_init_f)dlsym(hf, "AGR_Sate_Encoder_Init"); enc_encode = (enc_encode_f)dlsym(hf, "AGR_Sate_Encoder_Encode");
    enc_uninit = (enc_uninit_f)dlsym(hf, "AGR_Sate_Encoder_Uninit");
    dec_init = (dec_init_f)dlsym(hl, "AGR_Sate_Decoder_Init"); dec_decode = (dec_decode_f)dlsym(hl, "AGR_Sate_Decoder_Decode");
    dec_uninit = (dec_uninit_f)dlsym(hl, "AGR_Sate_Decoder_Uninit");
    FILE *f = fopen(argv[2], "rb"); if (!f) { perror("pcm"); return 1; }
    fseek(f, 0, SEEK_END); clip_n = ftell(f) / 2; fseek(f, 0, SEEK_SET);
    clip = malloc(clip_n * 2); if (fread(clip, 2, clip_n, f) != (size_t)clip_n) return 1; fclose(f);
    n_threads = atoi(argv[3]); streams_per_thread = atoi(argv[4]); packets = atoi(argv[5]);
    rate_bps = argc > 6 ? atoi(argv[6]) : 13600;
    pin_threads = argc > 7 ? atoi(argv[7]) : 0;
    {
        cpu_set_t set; CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0)
            for (int c = 0; c < CPU_SETSIZE && n_allowed < 1024; c++) if (CPU_ISSET(c, &set)) allowed[n_allowed++] = c;
    }
    if (n_threads < 1 || n_threads > 256) return 2;
    pthread_barrier_init(&bar, NULL, n_threads);
    pthread_t th[256];
    for (long t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker, (void *)t);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    double t0 = 1e300, t1 = 0; unsigned long long nb = 0, h = 0;
    for (int t = 0; t < n_threads; t++) { if (t_start[t] < t0) t0 = t_start[t]; if (t_end[t] > t1) t1 = t_end[t]; nb += bytes_out[t]; h ^= pcm_hash[t]; }
    double npk = (double)n_threads * streams_per_thread * packets;
    printf("{\"packets\": %.0f, \"seconds\": %.6f, \"packets_per_s\": %.2f, \"threads\": %d, \"streams\": %d, \"packets_per_stream\": %d, "
           "\"mean_payload_bytes\": %.3f, \"pcm_hash\": \"%016llx\", \"pinned\": %d, \"cpus_allowed\": %d, \"packets_per_s_per_thread\": %.1f}\n",
           npk, t1 - t0, npk / (t1 - t0), n_threads, n_threads * streams_per_thread, packets, nb / npk, h, pin_threads, n_allowed,
           npk / (t1 - t0) / n_threads);
    return 0;
}
