"""CPU: the product's host-side C (segmenter, synth, .testtensor loader, ABI surface) -- no GPU compute."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, Oracle

CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_segmenter_reproduces_reference_cli_text(path):
    g = np.load(path)
    p = g["out2"][:, 1]
    assert vadc_b200.segments_text(p) == str(g["stdout"])
    assert vadc_b200.segments_text(p, vadc_b200.seg_params(centiseconds=1)) == str(g["stdout_centi"])


def _rand_probs(seed, n):
    rng = np.random.default_rng(seed)
    # piecewise-constant-ish probabilities that cross both thresholds often
    p = np.clip(np.repeat(rng.uniform(0, 1, n // 3 + 1), 3)[:n] + rng.normal(0, 0.1, n), 0, 1)
    return p.astype(np.float32)


@pytest.mark.parametrize("seed", range(12))
def test_segmenter_equals_oracle_state_machine(seed):
    o = Oracle()
    n = [0, 1, 2, 3, 5, 17, 96, 97, 500, 1000, 2000, 4000][seed]
    p = _rand_probs(seed, n)
    for kw in ({}, {"threshold": 0.7}, {"min_silence_ms": 500.0, "min_speech_ms": 100.0}, {"speech_pad_ms": 200.0},
               {"neg_threshold_relative": 0.4}, {"centiseconds": 1}):
        assert vadc_b200.segments_text(p, vadc_b200.seg_params(**kw)) == o.segments_text(p, **kw)


def test_segmenter_streaming_equals_batch():
    p = _rand_probs(5, 3000)
    whole = vadc_b200.StreamSegmenter()
    segs = whole.feed(p) + whole.finish()
    rng = np.random.default_rng(0)
    pieces = vadc_b200.StreamSegmenter()
    got, i = [], 0
    while i < len(p):
        k = int(rng.integers(1, 200))
        got += pieces.feed(p[i:i + k])
        i += k
    got += pieces.finish()
    assert got == segs and len(segs) > 5
    text = "".join(whole.format(s) for s in segs)
    assert text == vadc_b200.segments_text(p)


def test_segmenter_edge_cases():
    # all speech: one segment closed at end of stream at chunk G-1 (vadc.c:1008-1021)
    p = np.ones(100, np.float32)
    assert vadc_b200.segments_text(p) == "0.00,%.2f\n" % (np.float32(99) * np.float32(1536 / 16000) + np.float32(0.03))
    # too short a tail is dropped: (G-1 - start) must exceed min_speech_chunks
    assert vadc_b200.segments_text(np.array([0, 0, 1, 1, 1, 1], np.float32)) == ""
    assert vadc_b200.segments_text(np.zeros(50, np.float32)) == ""
    assert vadc_b200.segments_text(np.zeros(0, np.float32)) == ""
    # the temp_end == 0 sentinel quirk: speech starting at chunk 0 whose silence begins at... chunk index 0 cannot be "set"
    p = np.array([1, 1, 1, 1, 0, 0, 0, 0, 0, 0], np.float32)
    assert vadc_b200.segments_text(p) == Oracle().segments_text(p)
    # probabilities between the thresholds keep the state
    p = np.array([0.6] + [0.4] * 20 + [0.1] * 5, np.float32)
    assert vadc_b200.segments_text(p) == Oracle().segments_text(p) != ""


def test_synth_is_deterministic_and_speechlike():
    a = vadc_b200.synth_pcm(5, 16000 * 20)
    b = vadc_b200.synth_pcm(5, 16000 * 20)
    c = vadc_b200.synth_pcm(6, 16000 * 20)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a).max() < 0.6 * 32768 and a.std() > 50
    assert not vadc_b200.synth_pcm(1, 1000, kind=1).any()
    w = vadc_b200.synth_pcm(1, 100000, kind=2)
    assert w.min() < -32000 and w.max() > 32000


def _declared_functions(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:silero_b200|vadc)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = vadc_b200.lib()
    names = []
    for h in ("silero_b200.h", "vadc_segmenter.h"):
        names += _declared_functions(os.path.join(ROOT, "include", h))
    assert len(names) > 30
    for n in names:
        assert hasattr(L, n), "libsilero_b200.so does not export " + n


def test_bad_weights_are_rejected_before_any_device_use():
    L = vadc_b200.lib()
    h = C.c_void_p()
    for blob in (b"", b"\x01\x00\x00\x00", b"\x02\x00\x00\x00\x01\x00\x00\x00", open(vadc_b200.WEIGHTS_PATH, "rb").read()[:5000]):
        rc = L.silero_b200_create(blob, C.c_size_t(len(blob)), None, C.byref(h))
        assert rc == -2 and not h.value, (rc, L.silero_b200_last_error())
    # a well-formed container with the wrong tensor count is not a v3.1 model
    from testtensor_io import dump_testtensor
    blob = dump_testtensor([("a", np.zeros((2, 2), np.float32))])
    assert L.silero_b200_create(blob, C.c_size_t(len(blob)), None, C.byref(h)) == -2
    assert b"99" in L.silero_b200_last_error()
    assert L.silero_b200_create(None, C.c_size_t(0), None, C.byref(h)) == -1
    # a header whose size arithmetic wraps in 32 bits (dims 32768 x 32768 -> size 2^30, size * 4 == 0 == nbytes): must be rejected, not read
    import struct
    evil = struct.pack("<ii", 1, 1) + struct.pack("<i", 1) + b"x" + struct.pack("<iiiii", 2, 32768, 32768, 1 << 30, 0)
    assert L.silero_b200_create(evil, C.c_size_t(len(evil)), None, C.byref(h)) == -2 and not h.value


def test_no_cpu_fallback_without_device():
    """On a machine without a GPU the engine must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vadc_b200.EngineError, match="no CPU fallback"):
        vadc_b200.Engine()


def test_product_does_not_touch_the_oracle():
    """Nothing under vadc_b200/ or include/ may reference oracle/ (it is test infrastructure)."""
    for base in ("vadc_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "silero_oracle" not in txt and "oracle/" not in txt and "libvadc_ref" not in txt, os.path.join(dirpath, f)


def test_python_constants_mirror_the_header():
    """vadc_b200/api.py restates the numerics knobs of include/silero_b200.h for the tests and the bench."""
    import re
    hdr = open(os.path.join(ROOT, "include", "silero_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define SILERO_B200_(\w+)\s+(\d+)\b", hdr)}
    for name in ("STFT_AUTO", "STFT_EXACT", "STFT_HYBRID", "LSTM_AUTO", "LSTM_FP32", "LSTM_TENSOR",
                 "LSTM_FAITHFUL", "LAYERS_AUTO", "LAYERS_FP32", "LAYERS_TENSOR", "LAYERS_FAITHFUL"):
        assert getattr(vadc_b200, name) == defs[name], name


def test_exact_kernels_contain_no_contracted_packed_arithmetic():
    """The exact path may not fuse a multiply with an add. Scalar mul.rn / add.rn are never contracted, but ptxas 12.9 turns
    mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with --fmad=false). The kernels therefore pack EITHER the multiplies (FMUL2 on the
    register pairs an LDS.128 delivers, feeding scalar adds -- the default) OR the adds (FADD2 fed by scalar multiplies), never both
    in one kernel. This reads the SASS of the built library: no FFMA2 in any exact-path kernel, no kernel with FMUL2 and FADD2,
    FMUL2 present in the STFT."""
    import re
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", vadc_b200.LIB_PATH], capture_output=True, text=True).stdout
    seen = {}
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0]
        if not any(k in name for k in ("stft_sym_kernel", "stft_logmag_kernel", "exact_front_kernel", "exact_layer_kernel", "exact_lstm_kernel",
                                        "faithful_encoder_kernel", "faithful_decoder_kernel")):
            continue
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", part, re.M)
        seen[name] = (ops.count("FFMA2"), ops.count("FMUL2"), ops.count("FADD2"))
    assert len(seen) >= 10, sorted(seen)
    for name, (ffma2, fmul2, fadd2) in seen.items():
        assert ffma2 == 0 and not (fmul2 > 0 and fadd2 > 0), (name, ffma2, fmul2, fadd2)
    assert any(v[1] > 0 for k, v in seen.items() if "stft_sym_kernel" in k)
