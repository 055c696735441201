"""Memory safety of the decoder arithmetic on malformed input: the host build of the kernel source under AddressSanitizer
and UBSan, fed bit-flipped, random, truncated and mislabelled payloads with arbitrary lost flags (tests/hostsim/fuzz_dec.cpp).
A decoder that reads outside its tables on a bad packet would, on the GPU, fault the kernel and take the whole batch down."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decoder_survives_malformed_packets_under_asan(tmp_path):
    exe = tmp_path / "fuzz_dec"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-w",
           os.path.join(ROOT, "tests", "hostsim", "fuzz_dec.cpp"), "-o", str(exe)]
    try:
        subprocess.check_call(cmd)
    except (subprocess.CalledProcessError, FileNotFoundError):
        pytest.skip("sanitizer build not available")
    p = subprocess.run([str(exe), "1500", "7"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "no memory errors" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
