"""CPU: pins the oracle restatement to the reference itself.

(1) against the committed golden vectors generated from the unmodified reference
    (tests/golden/e2e_*.npz, made by tests/golden/make_e2e_golden.py) -- always runs;
(2) against oracle/_ref built from /root/reference, bit for bit -- runs where that build exists.
"""
import glob
import os

import numpy as np
import pytest

from oracle_lib import ROOT, Oracle, Reference, have_ref, ref_cli

CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))


def test_golden_cases_exist():
    assert len(CASES) >= 4


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_restatement_matches_reference_golden_bitwise(path):
    g = np.load(path)
    o = Oracle()
    out = o.run_pcm(g["pcm"])
    assert out.shape == g["out2"].shape
    assert np.array_equal(out.view(np.uint32), g["out2"].view(np.uint32))
    o.reset()
    x = (g["pcm"][: 8 * 1536].astype(np.float32) / np.float32(32768.0)).reshape(8, 1536)
    st = o.run_stages(x)
    for k, v in st.items():
        assert np.array_equal(v.view(np.uint32), g["stages_" + k].view(np.uint32)), k
    assert o.segments_text(out[:, 1]) == str(g["stdout"])
    assert o.segments_text(out[:, 1], centiseconds=1) == str(g["stdout_centi"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_bit_identical_to_reference_build():
    import vadc_b200
    pcm = vadc_b200.synth_pcm(4242, 1536 * 64)
    x = (pcm.astype(np.float32) / np.float32(32768.0)).reshape(-1, 1536)
    o, r = Oracle(), Reference()
    a, b = o.run_stages(x), r.run_stages(x)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    # batch invariance of the reference (finding F6): batch 1 / 7 / 96 give identical bits
    r.reset()
    p96 = r.run_pcm(pcm, 96)
    r.reset()
    p7 = r.run_pcm(pcm, 7)
    r.reset()
    p1 = r.run_pcm(pcm, 1)
    assert np.array_equal(p96, p7) and np.array_equal(p96, p1)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_backend_init_contract():
    # silero.h:39-43
    assert Reference().config() == [-1, 0, 1536, 1536, 3]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("args", [(), ("--output_centi_seconds",), ("--threshold", "0.6", "--min_silence", "400"),
                                  ("--min_speech", "500", "--speech_pad", "100"), ("--neg_threshold_relative", "0.3")])
def test_segment_text_matches_reference_cli(args):
    import vadc_b200
    pcm = vadc_b200.synth_pcm(99, 16000 * 40 + 123)
    o = Oracle()
    p = o.run_pcm(pcm)[:, 1]
    kw = {}
    it = iter(args)
    for a in it:
        if a == "--output_centi_seconds":
            kw["centiseconds"] = 1
        else:
            kw[{"--threshold": "threshold", "--min_silence": "min_silence_ms", "--min_speech": "min_speech_ms",
                "--speech_pad": "speech_pad_ms", "--neg_threshold_relative": "neg_threshold_relative"}[a]] = float(next(it))
    assert o.segments_text(p, **kw) == ref_cli(pcm, *args)
