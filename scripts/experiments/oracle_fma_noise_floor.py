"""CPU: how far is the reference from ITSELF under a different compiler flag? Builds oracle/silero_oracle.c (the restatement that is
bit-identical to the pinned reference build, tests/test_oracle_vs_ref.py) once more with -mfma -ffp-contract=fast and compares the
per-chunk probabilities on long single streams. Measured (3000 chunks = 288 s per stream, seeds 4242 7 99 1234 31337):
max |dp| = 9.2e-5, 3.7e-4, 6.2e-4, 3.3e-4, 1.6e-4 -- the B200 engine stays within 4e-5 .. 9e-5 of the pinned build on the same
streams (scripts/gpu_krel_sweep.py). The 1e-4 bar is met against the PINNED build; it is tighter than the reference's own
reproducibility across compiler flags."""
import ctypes as C, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vadc_b200, oracle_lib
from oracle_lib import Oracle

so = os.path.join(tempfile.gettempdir(), "liboracle_fma.so")
subprocess.run(["gcc", "-O2", "-mavx2", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "oracle", "silero_oracle.c"), "-lm"], check=True)


class Contracted(Oracle):
    def __init__(self):
        self.lib = C.CDLL(so)
        self.lib.so_model_load_file.restype = C.c_void_p
        self.lib.so_segments_text.restype = C.c_size_t
        self.m = C.c_void_p(self.lib.so_model_load_file(os.fsencode(oracle_lib.WEIGHTS)))
        self.state = np.zeros(256, np.float32)


a, b = Oracle(), Contracted()
for seed in (4242, 7, 99, 1234, 31337):
    pcm = vadc_b200.synth_pcm(seed, 1536 * 3000)
    a.reset(); b.reset()
    print(seed, "max |dp| pinned vs FMA-contracted build of the same C code: %.2e" % float(np.abs(a.run_pcm(pcm) - b.run_pcm(pcm)).max()), flush=True)
