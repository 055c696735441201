#!/usr/bin/env python3
"""Tiny encode/decode run for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Covers 40 ms, 20 ms and joint-mode packets, odd batch sizes (shadowed lane group in the quantiser kernel) and lossy decode."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import solo_b200  # noqa: E402
from tests.util import load_clip, loss_flags  # noqa: E402

clip = load_clip()
for kw, spp in ((dict(), 640), (dict(framesize_ms=20), 320), (dict(joint_hb=1), 640)):
    N, T = 37, 3
    eb, db = solo_b200.EncoderBatch(N, **kw), solo_b200.DecoderBatch(N, **kw)
    for p in range(T):
        x = np.stack([clip[(s * 7919 + p) * spp % (len(clip) - spp):][:spp] for s in range(N)]).astype(np.int16)
        bits, nb = eb.encode(x, cap=128)
        flags = np.array([loss_flags(T, 40, seed=s + 1)[p] for s in range(N)], np.int32)
        flags[nb[:, 0] <= 0] = 1
        nb2 = nb.copy()
        b2 = bits.copy()
        for s in range(N):
            n0, n1 = int(nb[s, 0]), int(nb[s, 1])
            if flags[s] == 2:
                nb2[s] = (n0 - n1, 0)
            elif flags[s] == 3:
                b2[s, :n1] = bits[s, n0 - n1:n0]
                nb2[s] = (n1, 0)
            if nb2[s, 0] <= 0:
                nb2[s] = (16, 8)
        pcm, ret = db.decode(b2, nb2, flags)
        assert (ret == 0).all()
    eb.close(); db.close()
print("sanitize_small: done")
