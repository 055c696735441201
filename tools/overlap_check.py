#!/usr/bin/env python3
"""Experiment: does decode(t) on a second CUDA stream overlap usefully with encode(t+1)?

usage: python tools/overlap_check.py [streams] [steps]
Times the same K packet waves three ways (device-resident inputs, CUDA events, after warm-up):
  serial     encoder wave then decoder wave on one stream (what bench.py does)
  two-stream encoder on stream A, decoder on stream B, payload rows double-buffered, events between them
  two-stream-prio  the same with the decoder stream at high priority (its blocks are dispatched ahead of the encoder's)
Prints ms per wave of each and checks that both orders decode to the same PCM."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import solo_b200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W, CAP = 3, 128
T = K + W
dev = torch.device("cuda", 0)
clip = bench.load_clip()
d_pcm = torch.from_numpy(bench.speech_replay(clip, np.arange(N), T)).to(dev)
d_flags = torch.full((N,), 4, dtype=torch.int32, device=dev)


def run(mode):
    enc, dec = solo_b200.EncoderBatch(N, rate=bench.RATE), solo_b200.DecoderBatch(N)
    bits = [torch.zeros((N, CAP), dtype=torch.uint8, device=dev) for _ in range(2)]
    nb = [torch.zeros((N, 2), dtype=torch.int16, device=dev) for _ in range(2)]
    out = torch.zeros((N, 640), dtype=torch.int16, device=dev)
    ret = torch.zeros((N,), dtype=torch.int32, device=dev)
    sa, sb_ = torch.cuda.Stream(), torch.cuda.Stream(priority=-1 if mode == "two-stream-prio" else 0)
    enc_done = [torch.cuda.Event() for _ in range(2)]
    dec_done = [torch.cuda.Event() for _ in range(2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = torch.zeros((), dtype=torch.int64, device=dev)

    def step(t):
        b = t & 1
        if mode == "serial":
            enc.encode_device(d_pcm[t].data_ptr(), bits[b].data_ptr(), CAP, nb[b].data_ptr(), sa.cuda_stream)
            dec.decode_device(out.data_ptr(), bits[b].data_ptr(), CAP, nb[b].data_ptr(), d_flags.data_ptr(), ret.data_ptr(), sa.cuda_stream)
        else:
            sa.wait_event(dec_done[b])            # the decoder is done with this payload buffer (two waves ago)
            enc.encode_device(d_pcm[t].data_ptr(), bits[b].data_ptr(), CAP, nb[b].data_ptr(), sa.cuda_stream)
            enc_done[b].record(sa)
            sb_.wait_event(enc_done[b])
            dec.decode_device(out.data_ptr(), bits[b].data_ptr(), CAP, nb[b].data_ptr(), d_flags.data_ptr(), ret.data_ptr(), sb_.cuda_stream)
            dec_done[b].record(sb_)

    for t in range(W):
        step(t)
    torch.cuda.synchronize()
    e0.record(sa)
    for t in range(W, T):
        step(t)
    sa.wait_stream(sb_)
    e1.record(sa)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    chk = int(out.to(torch.int64).sum().item()), int(ret.abs().sum().item())
    enc.close(); dec.close()
    return ms, chk


for mode in ("serial", "two-stream-prio", "two-stream", "two-stream-prio"):
    ms, chk = run(mode)
    print("%-16s %7.3f ms/wave  %.3f M packets/s  pcm checksum %d ret %d" % (mode, ms, N / ms / 1e3, chk[0], chk[1]), flush=True)
