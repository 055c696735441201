"""The N > 1 path on CPU: world_size 2 (and 3, uneven split) over gloo.  Root scatters PCM rows, every rank encodes its
contiguous block of streams (here with the host build of the kernel source, since this container has no GPU), root
gathers payloads and length fields; the result must equal the single-process encoding stream for stream."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from solo_b200.shard import gather_streams, local_count, scatter_streams, shard_bounds
from tests.util import load_clip, speech_replay

CAP = 128


def test_shard_bounds_cover_every_stream_once():
    for n in (0, 1, 7, 4096, 65536, 524288):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
            assert sum(local_count(n, r, w) for r in range(w)) == n


def _encode_rows(encs, x):
    from tests.hostsim import sim
    bits = np.zeros((len(encs), CAP), np.uint8)
    nb = np.zeros((len(encs), 2), np.int16)
    for i, e in enumerate(encs):
        b, n2, n = e.encode(x[i])
        bits[i, :n] = np.frombuffer(b[:n], np.uint8)
        nb[i] = n2
    return bits, nb


def _worker(rank, world, port, n_streams, n_packets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.hostsim import sim
    pcm = torch.from_numpy(speech_replay(load_clip(), n_streams, n_packets)) if rank == 0 else None
    encs = [sim.SimEncoder(rate=13600) for _ in range(local_count(n_streams, rank, world))]
    out = []
    for p in range(n_packets):
        x = scatter_streams(pcm[p] if rank == 0 else None, n_streams, (640,), torch.int16, "cpu")
        bits, nb = _encode_rows(encs, x.numpy())
        fb = gather_streams(torch.from_numpy(bits), n_streams)
        fn = gather_streams(torch.from_numpy(nb), n_streams)
        if rank == 0:
            out.append((fb.numpy().copy(), fn.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(out)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_streams", [(2, 6), (3, 7)])
def test_scatter_encode_gather_equals_single_process(world, n_streams):
    from tests.hostsim import sim
    sim.lib()                           # build once before forking workers
    n_packets = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, n_packets, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    x = speech_replay(load_clip(), n_streams, n_packets)
    encs = [sim.SimEncoder(rate=13600) for _ in range(n_streams)]
    for p in range(n_packets):
        bits, nb = _encode_rows(encs, x[p])
        assert np.array_equal(got[p][1], nb)
        assert np.array_equal(got[p][0], bits)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_single_ingest_on_two_real_gpus(mode):
    """configs[3] plumbing with the real CUDA encoder on >= 2 GPUs of one node (skipped on a single-GPU box): rank 0 owns all
    PCM rows; `peer`: every rank encodes / decodes straight out of and into rank 0's HBM (CUDA IPC + NVLink, no transfer
    step), `nccl`: explicit scatter / gather.  tools/shard_check.py compares the gathered bytes with a single-GPU run."""
    import json
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541" if mode == "peer" else "29542", os.path.join(root, "tools", "shard_check.py"), "4096", "3", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-2000:]
    assert json.loads(lines[-1])["ok"] is True
