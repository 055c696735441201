"""The drop-in boundary without a GPU: libsolo_b200.so loads, exports every function include/*.h declares, compiles
against the headers from plain C, and refuses to work (NULL / -1 + an error text, never a CPU path) without a device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_functions():
    names = []
    for h in sorted(os.listdir(INC)):
        src = open(os.path.join(INC, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        names += re.findall(r"\b((?:AGR_Sate|solo_b200)_\w+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def so():
    from solo_b200 import build
    return build.build(verbose=False)


def test_headers_declare_the_reference_surface():
    f = declared_functions()
    for n in ("AGR_Sate_Encoder_Init", "AGR_Sate_Encoder_Encode", "AGR_Sate_Encoder_Uninit",
              "AGR_Sate_Decoder_Init", "AGR_Sate_Decoder_Decode", "AGR_Sate_Decoder_Uninit"):
        assert n in f
    assert len([n for n in f if n.startswith("solo_b200_")]) >= 12


def test_library_exports_every_declared_symbol(so):
    L = C.CDLL(so)
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_headers_compile_and_link_from_c(so, tmp_path):
    """A C translation unit written the way the reference's enc_main.c / dec_main.c use the API links against the
    library with nothing but the two headers (no C++, no CUDA, no torch types in the signatures)."""
    src = tmp_path / "client.c"
    src.write_text(r'''
#include "AGR_JC1_SDK_API.h"
#include "solo_b200.h"
#include <stdio.h>
int main(void) {
    USER_Ctrl_enc ec = {0, 13600, 16000, 0, 40, 0, 0, 0};
    USER_Ctrl_dec dc = {0, 16000, 40, 0, 0, 0};
    short pcm[640] = {0}, nb[6] = {0}, ns = 0;
    unsigned char bits[1024];
    void *e = AGR_Sate_Encoder_Init(&ec);
    void *d = AGR_Sate_Decoder_Init(&dc);
    if (!e || !d) { printf("init refused: %s\n", solo_b200_last_error()); }
    int r1 = AGR_Sate_Encoder_Encode(e, pcm, bits, 1024, nb);
    int r2 = AGR_Sate_Decoder_Decode(d, pcm, &ns, bits, nb, 4);
    int r3 = AGR_Sate_Encoder_Uninit(e), r4 = AGR_Sate_Decoder_Uninit(d);
    printf("%d %d %d %d state %d %d\n", r1, r2, r3, r4, solo_b200_enc_state_bytes(), solo_b200_dec_state_bytes());
    return (e && d) ? 0 : 3;
}
''')
    exe = tmp_path / "client"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INC, str(src), "-o", str(exe),
                           "-L", os.path.dirname(so), "-lsolo_b200", "-Wl,-rpath," + os.path.dirname(so)])
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert p.returncode == 0, p.stdout + p.stderr
    else:
        # no device: Init returns NULL with a reason, and the NULL-handle calls return -1 as the reference does
        # (AGR_BWE_SDK_API.c:139-141,160-162,261-263,289-291)
        assert p.returncode == 3, p.stdout + p.stderr
        assert "init refused:" in p.stdout and "-1 -1 -1 -1" in p.stdout


def test_no_cpu_fallback_when_library_missing(tmp_path):
    code = "import solo_b200, sys\ntry:\n    solo_b200.lib()\nexcept solo_b200.SoloError as e:\n    print('refused:', e); sys.exit(7)\n"
    env = dict(os.environ, SOLO_B200_LIB=str(tmp_path / "nope.so"), PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode == 7 and "no CPU fallback" in p.stdout, p.stdout + p.stderr


def test_batch_create_refuses_without_device(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import solo_b200
    with pytest.raises(solo_b200.SoloError):
        solo_b200.EncoderBatch(4, rate=13600, device=0)
    with pytest.raises(solo_b200.SoloError):
        solo_b200.DecoderBatch(4, device=0)


def test_product_does_not_reference_the_oracle():
    """oracle/ and tests/hostsim are checkers: nothing under solo_b200/ may import, link or open them."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "solo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                t = open(os.path.join(d, f), errors="ignore").read()
                if f.endswith(".py"):
                    t = re.sub(r'""".*?"""', "", t, flags=re.S)
                    t = re.sub(r"#[^\n]*", "", t)
                else:
                    t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
                    t = re.sub(r"//[^\n]*", "", t)
                if re.search(r"\boracle\b|hostsim|libjc1_|_ref/", t):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
