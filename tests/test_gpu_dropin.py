"""GPU: the drop-in boundary itself. oracle/_ref/vadc_b200_dropin is the reference's UNMODIFIED vadc.c
compiled against include/vadc_dropin/silero.h and linked to libsilero_b200.so (oracle/Makefile). Its
stdout must equal the reference CLI's, byte for byte, and its --raw_probabilities within 1e-4."""
import glob
import os
import subprocess

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, REF_CLI

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "oracle", "_ref", "vadc_b200_dropin")
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))


def run_cli(binary, pcm, *args):
    env = dict(os.environ, VADC_B200_WEIGHTS=vadc_b200.WEIGHTS_PATH)
    r = subprocess.run([binary, *args], input=np.ascontiguousarray(pcm, np.int16).tobytes(), capture_output=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-500:]
    return r.stdout.decode()


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="drop-in CLI not built (needs /root/reference at build time)")
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_reference_cli_on_b200_backend(path):
    g = np.load(path)
    assert run_cli(DROPIN, g["pcm"]) == str(g["stdout"])
    assert run_cli(DROPIN, g["pcm"], "--output_centi_seconds") == str(g["stdout_centi"])
    raw = np.array([float(v) for v in run_cli(DROPIN, g["pcm"], "--raw_probabilities").split()])
    assert raw.shape[0] == g["out2"].shape[0]
    assert np.abs(raw - g["out2"][:, 1]).max() <= 1e-4 + 1e-6
    # other batch sizes only regroup chunks (finding F6); 96 divides evenly into 8/16/32/48
    assert run_cli(DROPIN, g["pcm"], "--batch", "32") == str(g["stdout"])


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REF_CLI)), reason="needs both CLIs")
def test_60s_stream_same_stdout_as_reference_cli():
    pcm = vadc_b200.synth_pcm(31337, 16000 * 60)
    for args in ((), ("--threshold", "0.4"), ("--min_silence", "500", "--speech_pad", "60")):
        assert run_cli(DROPIN, pcm, *args) == run_cli(REF_CLI, pcm, *args)


def test_dropin_fails_cleanly_without_weights():
    if not os.path.exists(DROPIN):
        pytest.skip("drop-in CLI not built")
    r = subprocess.run([DROPIN, "--model", "/nonexistent.testtensor"], input=b"\0" * 4096, capture_output=True, timeout=60)
    assert b"silero_b200" in r.stderr and r.stdout == b""
