"""The faithful path's arithmetic (vadc_b200/csrc/faithful_kernel.cuh), compiled for the host and run serially, against the
oracle bit for bit: the normalization scalar, the four encoder layers and the decoder head. The LSTM's gate contraction is
checked below on the packed weight layout; the per-step orchestration of faithful_lstm_kernel by the GPU tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import ROOT, Oracle, Reference, _p, have_ref

import vadc_b200

SRC = os.path.join(ROOT, "tests", "hostcheck", "faithful_host_check.cpp")
SO = os.path.join(ROOT, "tests", "hostcheck", "_faithful_host.so")


@pytest.fixture(scope="module")
def host():
    subprocess.run(["g++", "-O2", "-mavx2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO, "-lm"], check=True)
    return C.CDLL(SO)


def tensor_table(o):
    o.lib.so_model_tensor.restype = C.c_void_p
    return (C.c_void_p * 99)(*[o.lib.so_model_tensor(o.m, i, None, None) for i in range(99)])


def signals():
    rng = np.random.default_rng(5)
    yield "speech", vadc_b200.synth_pcm(77, 1536 * 40).astype(np.float32) / np.float32(32768.0)
    yield "noise floor", (rng.standard_normal(1536 * 6) * 0.003).astype(np.float32)
    yield "full scale noise", rng.uniform(-1, 1, 1536 * 6).astype(np.float32)
    yield "zeros", np.zeros(1536 * 3, np.float32)


def test_encoder_and_decoder_bits(host):
    o = Oracle()
    tab = tensor_table(o)
    for name, x in signals():
        o.reset()
        st = o.run_stages(x)
        B = st["stft"].shape[0]
        a4 = np.zeros((B, 7, 64), np.float32)
        host.faithful_host_encoder(tab, _p(st["stft"]), B, _p(a4))
        want = np.ascontiguousarray(st["l4"].transpose(0, 2, 1))          # [B][64][7] -> token-major
        assert np.array_equal(a4.view(np.uint32), want.view(np.uint32)), (name, float(np.abs(a4 - want).max()))
        out = np.zeros((B, 2), np.float32)
        dw = np.ctypeslib.as_array(C.cast(tab[97], C.POINTER(C.c_float)), (128,))
        db = np.ctypeslib.as_array(C.cast(tab[98], C.POINTER(C.c_float)), (2,))
        host.faithful_host_decoder(_p(st["lstm"]), B, _p(dw), _p(db), _p(out))
        assert np.array_equal(out.view(np.uint32), st["out"].view(np.uint32)), (name, float(np.abs(out - st["out"]).max()))


def test_lstm_gate_order_bits(host):
    """fq::gate_dot on the kernel's packed weight layout, layer after layer as the kernels run, against the oracle's interleaved
    two-layer walk (lstm.c:156-218): outputs and final state bit for bit, state carried across two calls."""
    o = Oracle()
    tab = tensor_table(o)
    w = np.ctypeslib.as_array(C.cast(tab[95], C.POINTER(C.c_float)), (2, 256, 128))
    b = np.ctypeslib.as_array(C.cast(tab[96], C.POINTER(C.c_float)), (512,)).copy()
    wpack = np.ascontiguousarray(w.reshape(2, 256, 32, 4).transpose(0, 2, 1, 3))   # [layer][k/4][row][4]
    x = vadc_b200.synth_pcm(91, 1536 * 30).astype(np.float32) / np.float32(32768.0)
    h = np.zeros((2, 64), np.float32)
    c = np.zeros((2, 64), np.float32)
    for part in (x[:1536 * 18], x[1536 * 18:]):
        st = o.run_stages(part)                       # the oracle carries its own state across the two calls
        B = st["l4"].shape[0]
        seq = np.ascontiguousarray(st["l4"].transpose(0, 2, 1)).reshape(B * 7, 64)
        out = np.zeros((B * 7, 64), np.float32)
        host.faithful_host_lstm(_p(seq), B * 7, _p(h), _p(c), _p(wpack), _p(b), _p(out))
        assert np.array_equal(out.view(np.uint32), st["lstm"].reshape(B * 7, 64).view(np.uint32))
        assert np.array_equal(h.reshape(-1).view(np.uint32), o.state[:128].view(np.uint32))
        assert np.array_equal(c.reshape(-1).view(np.uint32), o.state[128:].view(np.uint32))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_encoder_bits_against_the_unmodified_reference_build(host):
    """Same comparison with the stage tensors of the UNMODIFIED reference backend (oracle/_ref/libvadc_ref.so) instead of the
    restatement: magnitudes in, fourth-layer output and probabilities (through the reference's own LSTM output) bit for bit."""
    o = Oracle()
    tab = tensor_table(o)
    x = vadc_b200.synth_pcm(123, 1536 * 24).astype(np.float32) / np.float32(32768.0)
    st = Reference().run_stages(x)
    B = st["stft"].shape[0]
    a4 = np.zeros((B, 7, 64), np.float32)
    host.faithful_host_encoder(tab, _p(st["stft"]), B, _p(a4))
    assert np.array_equal(a4.view(np.uint32), np.ascontiguousarray(st["l4"].transpose(0, 2, 1)).view(np.uint32))
    out = np.zeros((B, 2), np.float32)
    dw = np.ctypeslib.as_array(C.cast(tab[97], C.POINTER(C.c_float)), (128,))
    db = np.ctypeslib.as_array(C.cast(tab[98], C.POINTER(C.c_float)), (2,))
    host.faithful_host_decoder(_p(np.ascontiguousarray(st["lstm"])), B, _p(dw), _p(db), _p(out))
    assert np.array_equal(out.view(np.uint32), st["out"].view(np.uint32))


def test_encoder_orchestration_has_no_data_race_under_tsan():
    """tests/hostcheck/faithful_host_race.cpp: the encoder's stage loops run by host threads with fq::barrier() as a real barrier,
    under ThreadSanitizer: a missing or misplaced barrier is a reported data race (exit code 66; checked by deleting one), and the
    parallel result equals the serial one. Several thread counts, so that neighbouring elements land on different threads."""
    src = os.path.join(ROOT, "tests", "hostcheck", "faithful_host_race.cpp")
    exe = os.path.join(ROOT, "tests", "hostcheck", "_faithful_host_race")
    Oracle()                                                   # builds oracle/libsilero_oracle.so if it is not there yet
    r = subprocess.run(["g++", "-O1", "-g", "-mavx2", "-ffp-contract=off", "-fsanitize=thread", "-x", "c++", src, "-o", exe,
                        "-L" + os.path.join(ROOT, "oracle"), "-lsilero_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread", "-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in r.stderr.lower():
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr.strip().splitlines()[-1])
    assert r.returncode == 0, r.stderr
    for nt in (8, 5, 16):
        r = subprocess.run([exe, vadc_b200.WEIGHTS_PATH, str(nt)], capture_output=True, text=True)
        if "FATAL: ThreadSanitizer" in r.stderr:                # e.g. "unexpected memory mapping" under some kernels' ASLR settings
            pytest.skip("ThreadSanitizer cannot run here: " + r.stderr.strip().splitlines()[0])
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (nt, r.stdout, r.stderr[-2000:])
        assert "parallel == serial: yes" in r.stdout


def test_libm_restatements_against_the_c_library_over_their_whole_domains():
    """vadc_b200/csrc/libm_exact.cuh compiled for the host (the same source the kernels use) and swept against the C library of the
    pinned reference build: tanhf over every float with |x| <= 23 (+ the saturated range), expf over every float in [-105, 89]
    (overflow, the denormal results down to log(2^-150), underflow), log1pf over every non-negative float. The pytest run takes every
    3rd value (each run a different residue is not needed: the algorithms branch on exponent ranges, not on single mantissas);
    `tests/hostcheck/_libm_exhaustive 1` sweeps all 6.6e9 values in ~50 CPU-seconds (last full run: 0 / 2 / 0 mismatches -- the two
    expf arguments where the library's FMA build rounds its double polynomial the other way)."""
    src = os.path.join(ROOT, "tests", "hostcheck", "libm_exhaustive.cpp")
    exe = os.path.join(ROOT, "tests", "hostcheck", "_libm_exhaustive")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-I", os.path.join(ROOT, "vadc_b200", "csrc"), src, "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    last = r.stdout.strip().splitlines()[-1].split()
    assert last[0] == "RESULT" and int(last[2]) == 0 and int(last[4]) <= 2 and int(last[6]) == 0, r.stdout
