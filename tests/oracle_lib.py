"""ctypes access to the CPU oracles. TEST INFRASTRUCTURE ONLY.

  * oracle/libsilero_oracle.so   -- this repo's plain-C restatement (oracle/silero_oracle.c)
  * oracle/_ref/libvadc_ref.so   -- the unmodified reference C backend (built by oracle/Makefile
                                    where /root/reference exists; a prebuilt copy travels to the GPU box)
  * oracle/_ref/vadc_linux       -- the unmodified reference CLI
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libsilero_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvadc_ref.so")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "vadc_linux")
WEIGHTS = os.path.join(ROOT, "vadc_b200", "weights", "silero_v31_16k.testtensor")

STAGE_SHAPES = (("stft", (129, 25)), ("norm", (129, 25)), ("l1", (16, 13)), ("l2", (32, 7)), ("l3", (32, 7)),
                ("l4", (64, 7)), ("lstm", (7, 64)), ("out", (2,)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build_oracles():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


class SegParams(C.Structure):
    _fields_ = [("min_silence_ms", C.c_float), ("min_speech_ms", C.c_float), ("threshold", C.c_float),
                ("neg_threshold_relative", C.c_float), ("speech_pad_ms", C.c_float), ("centiseconds", C.c_int)]


class Oracle:
    """The restatement. Function names follow oracle/silero_oracle.h."""

    def __init__(self, weights=WEIGHTS):
        if not os.path.exists(ORACLE_SO):
            build_oracles()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.so_model_load_file.restype = C.c_void_p
        self.lib.so_model_load.restype = C.c_void_p
        self.lib.so_segments_text.restype = C.c_size_t
        if isinstance(weights, (bytes, bytearray)):
            self.m = C.c_void_p(self.lib.so_model_load(bytes(weights), C.c_size_t(len(weights))))
        else:
            self.m = C.c_void_p(self.lib.so_model_load_file(os.fsencode(weights)))
        assert self.m.value, "oracle could not load weights"
        self.state = np.zeros(256, np.float32)

    def reset(self):
        self.state[:] = 0

    def run_chunks(self, samples):
        x = np.ascontiguousarray(samples, np.float32).reshape(-1, 1536)
        out = np.zeros((x.shape[0], 2), np.float32)
        self.lib.so_run_chunks(self.m, _p(self.state), _p(x), x.shape[0], _p(out))
        return out

    def run_stages(self, samples):
        x = np.ascontiguousarray(samples, np.float32).reshape(-1, 1536)
        B = x.shape[0]
        o = {k: np.zeros((B,) + s, np.float32) for k, s in STAGE_SHAPES}
        self.lib.so_run_chunks_stages(self.m, _p(self.state), _p(x), B, *[_p(o[k]) for k, _ in STAGE_SHAPES])
        return o

    def run_pcm(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.zeros((len(pcm) // 1536, 2), np.float32)
        self.lib.so_run_pcm(self.m, _p(self.state), _p(pcm), C.c_longlong(len(pcm)), _p(out))
        return out

    def segments_text(self, prob, **kw):
        p = SegParams()
        self.lib.so_segment_params_default(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        prob = np.ascontiguousarray(prob, np.float32).reshape(-1)
        cap = 64 * (len(prob) + 2) + 64
        buf = C.create_string_buffer(cap)
        n = self.lib.so_segments_text(_p(prob), C.c_longlong(len(prob)), C.byref(p), buf, C.c_size_t(cap))
        return buf.raw[:n].decode()


def have_ref():
    return os.path.exists(REF_SO)


class Reference:
    """The unmodified reference backend (oracle/ref_harness.c)."""

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        self.lib.vadc_ref_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.vadc_ref_create())
        assert self.h.value

    def config(self):
        c = (C.c_int * 5)()
        self.lib.vadc_ref_config(self.h, c)
        return list(c)

    def reset(self):
        self.lib.vadc_ref_reset(self.h)

    def run_chunks(self, samples):
        x = np.ascontiguousarray(samples, np.float32).reshape(-1, 1536)
        out = np.zeros((x.shape[0], 2), np.float32)
        self.lib.vadc_ref_run(self.h, _p(x), x.shape[0], _p(out))
        return out

    def run_stages(self, samples):
        x = np.ascontiguousarray(samples, np.float32).reshape(-1, 1536)
        B = x.shape[0]
        o = {k: np.zeros((B,) + s, np.float32) for k, s in STAGE_SHAPES}
        self.lib.vadc_ref_stages(self.h, _p(x), B, *[_p(o[k]) for k, _ in STAGE_SHAPES])
        return o

    def run_pcm(self, pcm, batch=96):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.zeros((len(pcm) // 1536, 2), np.float32)
        self.lib.vadc_ref_run_pcm(self.h, _p(pcm), C.c_longlong(len(pcm)), batch, _p(out))
        return out


def ref_cli(pcm, *args):
    """stdout of the unmodified reference CLI fed with s16le on stdin."""
    r = subprocess.run([REF_CLI, *args], input=np.ascontiguousarray(pcm, np.int16).tobytes(), capture_output=True)
    return r.stdout.decode()
