#!/usr/bin/env python3
"""Development aid: cycle stamps of one block of the analysis kernel at every phase boundary.
Build with  python -m solo_b200.build -DSB_PHASE_TIMING -ovariants/lib_pt.so  and run
    SOLO_B200_LIB=variants/lib_pt.so python tools/phase_times.py [streams]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import solo_b200  # noqa: E402
from tests.util import load_clip, speech_replay  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
L = solo_b200.lib()
L.solo_b200_set_chunks(1)
x = speech_replay(load_clip(), N, 3, first_packet=20)
eb = solo_b200.EncoderBatch(N)
st = (C.c_longlong * 512)()
ln = (C.c_int * 512)()
for p in range(3):
    eb.encode(x[p], cap=128)
    n = L.sb_phase_times_read(st, ln, 512)
t = np.array(st[:n]); l = np.array(ln[:n])
src = open(os.path.join(os.path.dirname(__file__), "..", "solo_b200", "csrc", "sb_coop.cuh")).read().split("\n")
par = open(os.path.join(os.path.dirname(__file__), "..", "solo_b200", "csrc", "sb_par.cuh")).read().split("\n")
print("marks:", n, "total cycles:", int(t[-1] - t[0]))
agg = {}
for i in range(1, n):
    key = (int(l[i - 1]), int(l[i]))
    agg.setdefault(key, []).append(int(t[i] - t[i - 1]))
inst = sum(sum(v) for k, v in agg.items() if k[1] < 0)
print("inside instance sections: %d cycles (%.0f%%)" % (inst, 100.0 * inst / (t[-1] - t[0])))
for i in range(1, n):
    print("%4d -> %4d  %8d" % (l[i - 1], l[i], t[i] - t[i - 1]))
