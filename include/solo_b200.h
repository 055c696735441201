/*
 * solo_b200 -- batched extension of the SOLO SDK ABI (plain C, no CUDA or torch types in the signatures).
 *
 * One batch object owns N independent codec streams whose state stays resident in GPU memory between
 * calls; one call advances every stream by one 40 ms packet.  Semantics per stream are exactly those of
 * AGR_Sate_Encoder_Encode / AGR_Sate_Decoder_Decode (AGR_JC1_SDK_API.h; reference
 * /root/reference/JC1_SDK_SRC_ARM/src/libBWE/AGR_BWE_SDK_API.c:129-152, JC1_SDK_SRC_FLP/.../AGR_BWE_SDK_API.c:249-279).
 *
 * Two flavours of every call:
 *   *_host    : buffers are host memory (pinned or pageable); host<->device copies happen inside the call
 *   *_device  : buffers are device pointers; the call only enqueues kernels on `cuda_stream`
 *               (a cudaStream_t passed as void*; NULL = legacy default stream) and does not synchronise.
 *
 * Packet size: ctrl->framesize_ms = 40 (two 20 ms codec frames per packet, 8 high-band bytes, 640 samples per row) or 20
 * (one frame, 4 high-band bytes, 320 samples per row) -- the two packet sizes of the reference (AGR_BWE_SDK_API.c:78-110);
 * with joint_enable = 1, joint_mode = 1 (40 ms packets only) the high band is coded as one 40 ms frame (4 bytes, :63-66).
 *
 * Layouts (row = stream):  pcm  int16 [N][16 * framesize_ms]   bits uint8 [N][cap]
 *                          nbytes int16 [N][2]        ({total, len(MD2)+8} as the single-stream API)
 *                          lostflag int32 [N]         (1 lost, 2 MD1 only, 3 MD2+HB only, 4 both)
 * For decode, row i of `bits` holds the payload exactly as the caller would hand it to
 * AGR_Sate_Decoder_Decode (already trimmed for lostflag 2 / 3) and nbytes[i] the matching {n0, n1};
 * nbytes is NOT modified (the single-stream call rewrites its nBytes[]).
 * ret[i] receives the per-stream return code of the reference call (0 ok, negative SILK error codes).
 *
 * All functions return 0 on success, a negative value on error (-1 bad argument, -2 CUDA failure;
 * solo_b200_last_error() gives a text).  A missing GPU is an error: there is no CPU fallback.
 */
#ifndef SOLO_B200_H
#define SOLO_B200_H

#include <stdint.h>
#include "AGR_JC1_SDK_API.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct solo_b200_enc_batch solo_b200_enc_batch;
typedef struct solo_b200_dec_batch solo_b200_dec_batch;

/* device = CUDA device ordinal; ctrl is applied to every stream (same checks as AGR_Sate_Encoder_Init). */
solo_b200_enc_batch *solo_b200_enc_batch_create(int n_streams, const USER_Ctrl_enc *ctrl, int device);
int solo_b200_enc_batch_encode_host(solo_b200_enc_batch *b, const int16_t *pcm, uint8_t *bits, int cap, int16_t *nbytes);
int solo_b200_enc_batch_encode_device(solo_b200_enc_batch *b, const int16_t *d_pcm, uint8_t *d_bits, int cap,
                                      int16_t *d_nbytes, void *cuda_stream);
void solo_b200_enc_batch_destroy(solo_b200_enc_batch *b);

solo_b200_dec_batch *solo_b200_dec_batch_create(int n_streams, const USER_Ctrl_dec *ctrl, int device);
int solo_b200_dec_batch_decode_host(solo_b200_dec_batch *b, int16_t *pcm, const uint8_t *bits, int cap,
                                    const int16_t *nbytes, const int32_t *lostflag, int32_t *ret);
int solo_b200_dec_batch_decode_device(solo_b200_dec_batch *b, int16_t *d_pcm, const uint8_t *d_bits, int cap,
                                      const int16_t *d_nbytes, const int32_t *d_lostflag, int32_t *d_ret, void *cuda_stream);
void solo_b200_dec_batch_destroy(solo_b200_dec_batch *b);

/* Stream state hand-over (checkpoint, migration between batches / GPUs, streams joining and leaving a batch): the complete
   codec state of stream `idx` as an opaque blob of solo_b200_{enc,dec}_state_bytes() bytes.  Synchronous; a blob is only
   meaningful to the same build of the library. */
int solo_b200_enc_batch_export_state(solo_b200_enc_batch *b, int idx, void *blob);
int solo_b200_enc_batch_import_state(solo_b200_enc_batch *b, int idx, const void *blob);
int solo_b200_dec_batch_export_state(solo_b200_dec_batch *b, int idx, void *blob);
int solo_b200_dec_batch_import_state(solo_b200_dec_batch *b, int idx, const void *blob);

/* Packet framing around one payload row [MD1 | MD2 | HB] (reference drivers enc_main.c:212-234, dec_main.c:196-307, README
   "Bitstream sending" / "Bitstream screening and synthesis at receiver").  Host functions, no GPU involved.
     split : the two network packets of a row -- description 1 = MD1, description 2 = MD2 + 8 high-band bytes (pointers into
             `bits`); both empty for a DTX row (nbytes[0] == 0)
     merge : from whichever packets arrived (pass NULL / 0 for a missing one) build the decoder's row, nbytes[2] and lostflag
             (4 both, 2 only description 1, 3 only description 2, 1 none)
     bitfile_pack / unpack : one record of the reference's .bit file (int16 total, int16 len(MD2)+8, payload); return the
             record size, or -1
     apply_loss_device : the receiver-side trimming for a whole batch on the GPU (rows of `cap` bytes): lostflag 2 keeps MD1,
             3 moves MD2 + HB to the front, 4 / 1 pass through -- feeds solo_b200_dec_batch_decode_device directly */
int solo_b200_split_packet(const uint8_t *bits, const int16_t *nbytes, const uint8_t **p1, int *n1, const uint8_t **p2, int *n2);
int solo_b200_merge_packets(const uint8_t *p1, int n1, const uint8_t *p2, int n2, uint8_t *bits, int cap, int16_t *nbytes,
                            int32_t *lostflag);
int solo_b200_bitfile_pack(const uint8_t *bits, const int16_t *nbytes, uint8_t *out, int out_cap);
int solo_b200_bitfile_unpack(const uint8_t *in, int in_len, const uint8_t **payload, int16_t *nbytes);
int solo_b200_apply_loss_device(const uint8_t *d_bits_in, const int16_t *d_nbytes_in, const int32_t *d_lostflag, uint8_t *d_bits_out,
                                int16_t *d_nbytes_out, int cap, int n, void *cuda_stream);

/* Multi-GPU ingest through peer memory: after this call kernels launched on `device` may read and write buffers that live in
   `peer_device`'s memory (same process, or another process's allocation mapped with CUDA IPC).  The *_device entry points
   then accept such addresses for pcm / bits / nbytes / lostflag: the band-split kernel pulls its PCM rows over NVLink and
   the entropy-coding and decoder kernels push their result rows back, so a single-ingest deployment needs no separate
   scatter / gather step (solo_b200/shard.py shows the plumbing with torch.distributed). */
int solo_b200_enable_peer_access(int device, int peer_device);

/* The six AGR_Sate_* functions (AGR_JC1_SDK_API.h) hand out slots of a process-wide arena: segments of 256 streams that are
   created on demand and shared by the handles.  out[4] = {encoder segments, encoder slots in use, decoder segments, decoder
   slots in use}. */
int solo_b200_arena_stats(int out[4]);

/* Bytes of device state held per stream (encoder / decoder). */
int solo_b200_enc_state_bytes(void);
int solo_b200_dec_state_bytes(void);
/* Number of kernels this library has launched in this process (for benchmark bookkeeping). */
long long solo_b200_kernel_launches(void);
/* Kernel timing with CUDA events recorded on the launching stream (only while enabled).  profile_read fills two arrays
   of 4: total ms and launch count since the last read for {encoder analysis, encoder NSQ, encoder finish, decode}. */
void solo_b200_profile_enable(int on);
int solo_b200_profile_read(double ms_total[4], long long launches[4]);
/* A packet wave can be processed as `chunks` groups of streams on internal CUDA streams so that the copies of one group
   overlap the kernels of another.  Default (0, or env SOLO_B200_CHUNKS unset): 3 for the *_host entry points, 1 (one launch
   per kernel on the caller's stream) for the *_device entry points.  Results do not depend on it. */
void solo_b200_set_chunks(int chunks);
const char *solo_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
