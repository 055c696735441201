#!/usr/bin/env python3
"""Multi-GPU ingest through one root, on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py [streams] [packets]

Rank 0 holds the PCM of all streams on its GPU, scatters contiguous blocks of rows to the ranks with grouped NCCL
send/recv (solo_b200/shard.py), every rank encodes its block, rank 0 gathers payloads + length fields and compares them
byte for byte with the same streams encoded on GPU 0 alone.  Prints one JSON line with the scatter / gather times."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import solo_b200  # noqa: E402
from solo_b200.shard import gather_streams, local_count, scatter_streams, shard_bounds  # noqa: E402
from tests.util import load_clip, speech_replay  # noqa: E402

CAP = 128


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    mode = sys.argv[3] if len(sys.argv) > 3 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if mode == "peer":
        return main_peer(N, T, rank, world, local, dev)
    n_loc = local_count(N, rank, world)
    enc = solo_b200.EncoderBatch(n_loc, device=local)
    d_bits = torch.zeros((n_loc, CAP), dtype=torch.uint8, device=dev)
    d_nb = torch.zeros((n_loc, 2), dtype=torch.int16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    pcm_all = torch.from_numpy(speech_replay(load_clip(), N, T)).to(dev) if rank == 0 else None
    got, t_sc, t_ga = [], 0.0, 0.0
    for p in range(T):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        x = scatter_streams(pcm_all[p] if rank == 0 else None, N, (640,), torch.int16, dev)
        torch.cuda.synchronize(); t_sc += (time.perf_counter() - t0) if p else 0.0     # wave 0 sets up the NCCL channels
        enc.encode_device(x.data_ptr(), d_bits.data_ptr(), CAP, d_nb.data_ptr(), st)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        fb = gather_streams(d_bits, N)
        fn = gather_streams(d_nb, N)
        torch.cuda.synchronize(); t_ga += (time.perf_counter() - t0) if p else 0.0
        if rank == 0:
            got.append((fb.cpu().numpy().copy(), fn.cpu().numpy().copy()))
    enc.close()
    ok = True
    if rank == 0:
        ref = solo_b200.EncoderBatch(N, device=local)
        rb = torch.zeros((N, CAP), dtype=torch.uint8, device=dev)
        rn = torch.zeros((N, 2), dtype=torch.int16, device=dev)
        for p in range(T):
            ref.encode_device(pcm_all[p].data_ptr(), rb.data_ptr(), CAP, rn.data_ptr(), st)
            torch.cuda.synchronize()
            ok = ok and np.array_equal(got[p][1], rn.cpu().numpy()) and np.array_equal(got[p][0], rb.cpu().numpy())
        ref.close()
        print(json.dumps({"check": "nccl scatter -> encode on %d GPUs -> gather == single-GPU encode" % world, "ok": bool(ok),
                          "streams": N, "packets": T, "bounds": shard_bounds(N, world),
                          "scatter_ms_per_wave": 1e3 * t_sc / max(T - 1, 1), "gather_ms_per_wave": 1e3 * t_ga / max(T - 1, 1),
                          "scatter_GBps": N * 1280 * (world - 1) / world / max(t_sc / max(T - 1, 1), 1e-9) / 1e9}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def main_peer(N, T, rank, world, local, dev):
    """No transfer step: every rank encodes (and decodes) its rows straight out of / into the root's HBM."""
    from solo_b200.shard import share_from_root
    lo, hi = shard_bounds(N, world)[rank]
    n_loc = hi - lo
    if rank == 0:
        pcm_all = torch.from_numpy(speech_replay(load_clip(), N, T)).to(dev)
        bits_all = torch.zeros((N, CAP), dtype=torch.uint8, device=dev)
        nb_all = torch.zeros((N, 2), dtype=torch.int16, device=dev)
        out_all = torch.zeros((N, 640), dtype=torch.int16, device=dev)
    else:
        pcm_all = bits_all = nb_all = out_all = None
    pcm_all, bits_all, nb_all, out_all = (share_from_root(t) for t in (pcm_all, bits_all, nb_all, out_all))
    enc = solo_b200.EncoderBatch(n_loc, device=local)
    dec = solo_b200.DecoderBatch(n_loc, device=local)
    flags = torch.full((n_loc,), 4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    # the same waves with rank-local buffers, for the time comparison
    loc_pcm = pcm_all[:, lo:hi].to(dev).contiguous()
    loc_bits = torch.zeros((n_loc, CAP), dtype=torch.uint8, device=dev)
    loc_nb = torch.zeros((n_loc, 2), dtype=torch.int16, device=dev)
    loc_out = torch.zeros((n_loc, 640), dtype=torch.int16, device=dev)
    enc2 = solo_b200.EncoderBatch(n_loc, device=local)
    dec2 = solo_b200.DecoderBatch(n_loc, device=local)

    def wave(e, d, pcm_row_ptr, bits, nb, out):
        e.encode_device(pcm_row_ptr, bits.data_ptr(), CAP, nb.data_ptr(), st)
        d.decode_device(out.data_ptr(), bits.data_ptr(), CAP, nb.data_ptr(), flags.data_ptr(), 0, st)

    got, t_peer, t_loc = [], 0.0, 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for p in range(T):
        torch.cuda.synchronize(); dist.barrier()
        ev[0].record()
        wave(enc, dec, pcm_all[p, lo:hi].data_ptr(), bits_all[lo:hi], nb_all[lo:hi], out_all[lo:hi])
        ev[1].record()
        torch.cuda.synchronize(); dist.barrier()
        ev[2].record()
        wave(enc2, dec2, loc_pcm[p].data_ptr(), loc_bits, loc_nb, loc_out)
        ev[3].record()
        torch.cuda.synchronize(); dist.barrier()
        if p:
            t_peer += ev[0].elapsed_time(ev[1]); t_loc += ev[2].elapsed_time(ev[3])
        if rank == 0:
            got.append((bits_all.cpu().numpy().copy(), nb_all.cpu().numpy().copy(), out_all.cpu().numpy().copy()))
    tt = torch.tensor([t_peer, t_loc], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        ref, rdec = solo_b200.EncoderBatch(N, device=local), solo_b200.DecoderBatch(N, device=local)
        rb = torch.zeros((N, CAP), dtype=torch.uint8, device=dev)
        rn = torch.zeros((N, 2), dtype=torch.int16, device=dev)
        ro = torch.zeros((N, 640), dtype=torch.int16, device=dev)
        rf = torch.full((N,), 4, dtype=torch.int32, device=dev)
        for p in range(T):
            ref.encode_device(pcm_all[p].data_ptr(), rb.data_ptr(), CAP, rn.data_ptr(), st)
            rdec.decode_device(ro.data_ptr(), rb.data_ptr(), CAP, rn.data_ptr(), rf.data_ptr(), 0, st)
            torch.cuda.synchronize()
            ok = ok and np.array_equal(got[p][1], rn.cpu().numpy()) and np.array_equal(got[p][0], rb.cpu().numpy()) and np.array_equal(got[p][2], ro.cpu().numpy())
        bytes_moved = N * (world - 1) / world * (1280 + CAP + 4 + CAP + 4 + 4 + 1280)
        print(json.dumps({"check": "peer-memory ingest: %d GPUs encode + decode out of / into rank 0's HBM == single-GPU run" % world,
                          "ok": bool(ok), "streams": N, "packets": T, "bounds": shard_bounds(N, world),
                          "wave_ms_peer": float(tt[0]) / max(T - 1, 1), "wave_ms_local": float(tt[1]) / max(T - 1, 1),
                          "nvlink_bytes_per_wave": bytes_moved}))
    dist.barrier()
    for o in (enc, dec, enc2, dec2):
        o.close()
    del pcm_all, bits_all, nb_all, out_all
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
