"""Generates tests/golden/e2e_*.npz from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) so that parity stays pinned on machines where the reference
sources do not exist (the GPU box).

Each file holds: pcm (int16 s16le stream), out2 ([N,2] per-chunk outputs of the reference backend,
batch 96), stdout / stdout_centi (the reference CLI's segment text), stages_* (per-stage tensors
of the first 8 chunks). Run from the repo root:  python tests/golden/make_e2e_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vadc_b200  # synth + segmenter are host C; no GPU needed here
from oracle_lib import Reference, have_ref, ref_cli

CASES = {
    # name: (seed, kind, seconds)
    "bursts_a": (11, 0, 30),
    "bursts_b": (12, 0, 30),
    "zeros": (13, 1, 6),
    "white": (14, 2, 6),
}


def main():
    assert have_ref(), "build oracle/_ref first: make -C oracle ref"
    ref = Reference()
    for name, (seed, kind, secs) in CASES.items():
        n = 16000 * secs + 700  # a partial trailing chunk that must be dropped (vadc.c:964)
        pcm = vadc_b200.synth_pcm(seed, n, kind)
        ref.reset()
        out2 = ref.run_pcm(pcm)
        ref.reset()
        x = (pcm[: 8 * 1536].astype(np.float32) / np.float32(32768.0)).reshape(8, 1536)
        st = ref.run_stages(x)
        stdout = ref_cli(pcm)
        stdout_centi = ref_cli(pcm, "--output_centi_seconds")
        raw = ref_cli(pcm, "--raw_probabilities")
        cli_probs = np.array([float(v) for v in raw.split()], np.float64)
        assert len(cli_probs) == out2.shape[0]
        assert np.abs(cli_probs - out2[:, 1]).max() < 1e-6  # "%f" prints 6 decimals
        path = os.path.join(ROOT, "tests", "golden", "e2e_%s.npz" % name)
        np.savez_compressed(path, pcm=pcm, out2=out2, stdout=np.array(stdout), stdout_centi=np.array(stdout_centi),
                            **{"stages_" + k: v for k, v in st.items()})
        print(name, "chunks", out2.shape[0], "speech frac %.2f" % (out2[:, 1] >= 0.5).mean(), "segments", stdout.count("\n"),
              "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
