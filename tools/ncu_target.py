#!/usr/bin/env python3
"""Small encode(+decode) run to capture under ncu: N streams x T packet waves of speech-replay input, one pipeline chunk.
    ncu --set full --import-source on -k regex:sb_enc_analysis -c 1 -o out python tools/ncu_target.py 8192 3"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import solo_b200  # noqa: E402
from tests.util import load_clip, speech_replay  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dec = len(sys.argv) > 3
solo_b200.lib().solo_b200_set_chunks(1)
x = speech_replay(load_clip(), N, T, first_packet=20)
eb = solo_b200.EncoderBatch(N)
db = solo_b200.DecoderBatch(N) if dec else None
for p in range(T):
    bits, nb = eb.encode(x[p], cap=128)
    if db:
        db.decode(bits, nb, np.full(N, 4, np.int32))
print("ncu_target: done", N, T)
