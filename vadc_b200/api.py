"""ctypes binding of libsilero_b200.so (include/silero_b200.h, include/vadc_segmenter.h).

This is the thin Python face of the C ABI used by tests/ and bench.py. The product is the shared
library: hand-written sm_100a CUDA kernels plus C host code. There is no CPU fallback anywhere in
this package: if the library is missing, or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsilero_b200.so")
WEIGHTS_PATH = os.path.join(_HERE, "weights", "silero_v31_16k.testtensor")

CHUNK = 1536
SAMPLE_RATE = 16000
STFT_AUTO, STFT_EXACT, STFT_HYBRID = 0, 1, 4
LSTM_AUTO, LSTM_FP32, LSTM_TENSOR, LSTM_FAITHFUL = 0, 1, 2, 3
LAYERS_AUTO, LAYERS_FP32, LAYERS_TENSOR, LAYERS_FAITHFUL = 0, 1, 2, 3


class EngineError(RuntimeError):
    pass


class Opts(C.Structure):
    _fields_ = [("device", C.c_int), ("max_streams", C.c_int), ("window_chunks", C.c_int), ("stft_mode", C.c_int),
                ("stft_k_rel", C.c_float), ("lstm_mode", C.c_int), ("layer_mode", C.c_int), ("reserved", C.c_int * 1)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("batch_size_restriction", "is_silero_v5", "input_size_min", "input_size_max",
                                        "output_dims", "sm_count", "max_streams", "window_chunks")]


class SegParams(C.Structure):
    _fields_ = [("min_silence_ms", C.c_float), ("min_speech_ms", C.c_float), ("threshold", C.c_float),
                ("neg_threshold_relative", C.c_float), ("speech_pad_ms", C.c_float), ("chunk_samples", C.c_int),
                ("centiseconds", C.c_int)]


class Segment(C.Structure):
    _fields_ = [("start_chunk", C.c_int), ("end_chunk", C.c_int)]


class Segmenter(C.Structure):
    _fields_ = [("p", SegParams), ("min_speech_chunks", C.c_int), ("min_silence_chunks", C.c_int),
                ("neg_threshold", C.c_float), ("seconds_per_chunk", C.c_float), ("temp_end", C.c_int),
                ("current_speech_start", C.c_int), ("triggered", C.c_int), ("buffered", Segment),
                ("buffered_valid", C.c_int), ("global_chunk_index", C.c_int)]


_lib = None


def lib():
    """Loads the in-tree shared library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no fallback implementation)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.silero_b200_last_error.restype = C.c_char_p
        _lib.vadc_segments_text.restype = C.c_size_t
        _lib.vadc_segmenter_feed.restype = C.c_longlong
        _lib.vadc_segmenter_finish.restype = C.c_longlong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Engine:
    """One engine per GPU (silero_b200 handle)."""

    def __init__(self, weights=None, device=0, max_streams=1, window_chunks=0, stft_mode=0, stft_k_rel=0.0, lstm_mode=0, layer_mode=0):
        L = lib()
        opts = Opts()
        L.silero_b200_default_opts(C.byref(opts))
        opts.device, opts.max_streams, opts.window_chunks = device, max_streams, window_chunks
        opts.stft_mode, opts.stft_k_rel, opts.lstm_mode, opts.layer_mode = stft_mode, stft_k_rel, lstm_mode, layer_mode
        self._h = C.c_void_p()
        if weights is None:
            weights = WEIGHTS_PATH
        if isinstance(weights, (bytes, bytearray)):
            rc = L.silero_b200_create(bytes(weights), C.c_size_t(len(weights)), C.byref(opts), C.byref(self._h))
        else:
            rc = L.silero_b200_create_from_file(os.fsencode(weights), C.byref(opts), C.byref(self._h))
        self._check(rc)
        self.max_streams = max_streams

    def _check(self, rc):
        if rc != 0:
            raise EngineError("silero_b200 error %d: %s" % (rc, lib().silero_b200_last_error().decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().silero_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        i = Info()
        self._check(lib().silero_b200_get_info(self._h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in Info._fields_}

    # ---- backend_run semantics -------------------------------------------------------------
    def run_chunks(self, samples, stream=0):
        x = _f32(samples).reshape(-1, CHUNK)
        out = np.zeros((x.shape[0], 2), np.float32)
        self._check(lib().silero_b200_run_chunks(self._h, stream, _p(x), x.shape[0], _p(out)))
        return out

    # ---- multi-stream ------------------------------------------------------------------------
    def run_streams(self, pcm, nchunks=None, first_stream=0, want_out2=False):
        """pcm: int16 [S, nsamples] host array (C-contiguous rows). Returns probs [S, nchunks] (and out2)."""
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.strides[1] == 2
        S = pcm.shape[0]
        if nchunks is None:
            nchunks = pcm.shape[1] // CHUNK
        probs = np.zeros((S, nchunks), np.float32)
        out2 = np.zeros((S, nchunks, 2), np.float32) if want_out2 else None
        self._check(lib().silero_b200_run_streams(self._h, _p(pcm), C.c_longlong(pcm.strides[0] // 2), first_stream, S, nchunks,
                                                  _p(probs), _p(out2)))
        return (probs, out2) if want_out2 else probs

    def run_streams_ptr(self, pcm_ptr, stream_stride, nstreams, nchunks, probs_ptr, out2_ptr=None, first_stream=0):
        self._check(lib().silero_b200_run_streams(self._h, C.c_void_p(pcm_ptr), C.c_longlong(stream_stride), first_stream, nstreams,
                                                  nchunks, C.c_void_p(probs_ptr), C.c_void_p(out2_ptr) if out2_ptr else None))

    def run_streams_device(self, d_pcm, stream_stride, nstreams, nchunks, d_probs, d_out2=None, first_stream=0):
        self._check(lib().silero_b200_run_streams_device(self._h, C.c_void_p(d_pcm), C.c_longlong(stream_stride), first_stream,
                                                         nstreams, nchunks, C.c_void_p(d_probs) if d_probs else None,
                                                         C.c_void_p(d_out2) if d_out2 else None))

    def sync(self):
        self._check(lib().silero_b200_sync(self._h))

    # ---- on-device segmenter --------------------------------------------------------------------
    def segments_configure(self, params=None):
        self._check(lib().silero_b200_segments_configure(self._h, C.byref(params) if params is not None else None))

    def segments_reset(self, first_stream=0, nstreams=None):
        self._check(lib().silero_b200_segments_reset(self._h, first_stream, self.max_streams - first_stream if nstreams is None else nstreams))

    def run_streams_segments(self, pcm, nchunks=None, end_of_stream=False, cap=None, first_stream=0, want_probs=False):
        """pcm: int16 [S, nsamples] host array (or None with end_of_stream to only flush). Returns a list of
        [(start_chunk, end_chunk), ...] per stream (pairs finished by this call), plus probs if asked."""
        if pcm is None:
            S, nchunks = self.max_streams - first_stream, 0
            ptr, stride = None, 0
        else:
            assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.strides[1] == 2
            S = pcm.shape[0]
            if nchunks is None:
                nchunks = pcm.shape[1] // CHUNK
            ptr, stride = _p(pcm), pcm.strides[0] // 2
        if cap is None:
            cap = nchunks // 2 + 2
        segs = np.zeros((S, cap, 2), np.int32)
        counts = np.zeros(S, np.int32)
        probs = np.zeros((S, nchunks), np.float32) if want_probs else None
        self._check(lib().silero_b200_run_streams_segments(self._h, ptr, C.c_longlong(stride), first_stream, S, nchunks,
                                                           1 if end_of_stream else 0, _p(segs), cap, _p(counts), _p(probs)))
        out = [[(int(a), int(b)) for a, b in segs[s, :min(int(counts[s]), cap)]] for s in range(S)]
        return (out, counts, probs) if want_probs else (out, counts)

    def run_streams_segments_ptr(self, pcm_ptr, stream_stride, nstreams, nchunks, end_of_stream, segs_ptr, cap, counts_ptr, probs_ptr=None, first_stream=0):
        """Raw-pointer form (pinned host buffers): segs int32 [nstreams][cap][2], counts int32 [nstreams]."""
        self._check(lib().silero_b200_run_streams_segments(self._h, C.c_void_p(pcm_ptr) if pcm_ptr else None, C.c_longlong(stream_stride), first_stream,
                                                           nstreams, nchunks, 1 if end_of_stream else 0, C.c_void_p(segs_ptr), cap,
                                                           C.c_void_p(counts_ptr), C.c_void_p(probs_ptr) if probs_ptr else None))

    def submit_streams_segments_ptr(self, pcm_ptr, stream_stride, nstreams, nchunks, end_of_stream, segs_ptr, cap, counts_ptr, probs_ptr=None, first_stream=0):
        """Asynchronous run_streams_segments_ptr: returns a ticket for wait()."""
        t = C.c_ulonglong()
        self._check(lib().silero_b200_submit_streams_segments(self._h, C.c_void_p(pcm_ptr) if pcm_ptr else None, C.c_longlong(stream_stride), first_stream,
                                                              nstreams, nchunks, 1 if end_of_stream else 0, C.c_void_p(segs_ptr) if segs_ptr else None, cap,
                                                              C.c_void_p(counts_ptr) if counts_ptr else None, C.c_void_p(probs_ptr) if probs_ptr else None,
                                                              C.byref(t)))
        return int(t.value)

    def wait(self, ticket):
        self._check(lib().silero_b200_wait(self._h, C.c_ulonglong(ticket)))

    def run_streams_segments_device(self, d_pcm, stream_stride, nstreams, nchunks, end_of_stream, d_segs, cap, d_counts, d_probs=None, first_stream=0):
        self._check(lib().silero_b200_run_streams_segments_device(self._h, C.c_void_p(d_pcm) if d_pcm else None, C.c_longlong(stream_stride), first_stream,
                                                                  nstreams, nchunks, 1 if end_of_stream else 0,
                                                                  C.c_void_p(d_probs) if d_probs else None, C.c_void_p(d_segs), cap, C.c_void_p(d_counts)))

    def segment_probs_device(self, d_probs, stride, nstreams, nchunks, end_of_stream, d_segs, cap, d_counts, first_stream=0):
        self._check(lib().silero_b200_segment_probs_device(self._h, C.c_void_p(d_probs) if d_probs else None, C.c_longlong(stride), first_stream, nstreams,
                                                           nchunks, 1 if end_of_stream else 0, C.c_void_p(d_segs), cap, C.c_void_p(d_counts)))

    def reset(self, first_stream=0, nstreams=None):
        self._check(lib().silero_b200_reset(self._h, first_stream, self.max_streams - first_stream if nstreams is None else nstreams))

    def get_state(self, stream=0):
        h = np.zeros(128, np.float32)
        c = np.zeros(128, np.float32)
        self._check(lib().silero_b200_get_state(self._h, stream, _p(h), _p(c)))
        return h.reshape(2, 64), c.reshape(2, 64)

    def set_state(self, h, c, stream=0):
        h, c = _f32(h).reshape(128), _f32(c).reshape(128)
        self._check(lib().silero_b200_set_state(self._h, stream, _p(h), _p(c)))

    # ---- device helpers -------------------------------------------------------------------------
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(lib().silero_b200_device_alloc(self._h, C.c_size_t(nbytes), C.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._check(lib().silero_b200_device_free(self._h, C.c_void_p(ptr)))

    def h2d(self, d_ptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(lib().silero_b200_memcpy_h2d(self._h, C.c_void_p(d_ptr), _p(arr), C.c_size_t(arr.nbytes)))

    def d2h(self, arr, d_ptr):
        assert arr.flags["C_CONTIGUOUS"]
        self._check(lib().silero_b200_memcpy_d2h(self._h, _p(arr), C.c_void_p(d_ptr), C.c_size_t(arr.nbytes)))

    def stft_stats(self, reset=False):
        tot, ex = C.c_ulonglong(), C.c_ulonglong()
        self._check(lib().silero_b200_stft_stats(self._h, C.byref(tot), C.byref(ex), 1 if reset else 0))
        return int(tot.value), int(ex.value)

    def stage_libm(self, x):
        x = _f32(x).reshape(-1)
        e, t, l = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        self._check(lib().silero_b200_stage_libm(self._h, _p(x), x.size, _p(e), _p(t), _p(l)))
        return e, t, l

    def debug_wavefront(self, stall_producer=False, spin_limit=0):
        self._check(lib().silero_b200_debug_wavefront(self._h, int(stall_producer), int(spin_limit)))

    def set_profiling(self, on):
        self._check(lib().silero_b200_set_profiling(self._h, 1 if on else 0))

    def last_timing(self):
        ms = (C.c_float * 8)()
        n = C.c_longlong()
        self._check(lib().silero_b200_last_timing(self._h, ms, C.byref(n)))
        names = ("total", "stft", "layer1", "layer2", "layer3", "layer4", "lstm0", "lstm1_decoder")
        return dict(zip(names, [float(v) for v in ms])), int(n.value)

    def timer_start(self):
        self._check(lib().silero_b200_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(lib().silero_b200_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def measure_fp32_peak(self):
        tf = C.c_float()
        self._check(lib().silero_b200_measure_fp32_peak(self._h, C.byref(tf)))
        return float(tf.value)

    # ---- parity taps ----------------------------------------------------------------------------
    def measure_fp32_unfused_peak(self):
        v = C.c_float()
        self._check(lib().silero_b200_measure_fp32_unfused_peak(self._h, C.byref(v)))
        return float(v.value)

    def stage_stft_magnitude(self, samples):
        x = _f32(samples).reshape(-1, CHUNK)
        out = np.zeros((x.shape[0], 129, 25), np.float32)
        self._check(lib().silero_b200_stage_stft_magnitude(self._h, _p(x), x.shape[0], _p(out)))
        return out

    def stage_stft_norm(self, samples):
        x = _f32(samples).reshape(-1, CHUNK)
        norm = np.zeros((x.shape[0], 129, 25), np.float32)
        logmag = np.zeros((x.shape[0], 129, 25), np.float32)
        self._check(lib().silero_b200_stage_stft_norm(self._h, _p(x), x.shape[0], _p(norm), _p(logmag)))
        return norm, logmag

    def stage_norm(self, magnitude):
        m = _f32(magnitude).reshape(-1, 129, 25)
        out = np.zeros_like(m)
        self._check(lib().silero_b200_stage_norm(self._h, _p(m), m.shape[0], _p(out)))
        return out

    def stage_encoder(self, norm):
        x = _f32(norm).reshape(-1, 129, 25)
        B = x.shape[0]
        outs = [np.zeros(s, np.float32) for s in ((B, 16, 13), (B, 32, 7), (B, 32, 7), (B, 64, 7))]
        self._check(lib().silero_b200_stage_encoder(self._h, _p(x), B, *[_p(o) for o in outs]))
        return outs

    def stage_pipeline(self, samples):
        x = _f32(samples).reshape(-1, CHUNK)
        B = x.shape[0]
        outs = [np.zeros(s, np.float32) for s in ((B, 16, 13), (B, 32, 7), (B, 32, 7), (B, 64, 7))]
        self._check(lib().silero_b200_stage_pipeline(self._h, _p(x), B, *[_p(o) for o in outs]))
        return outs

    def stage_exact_pipeline(self, samples):
        x = _f32(samples).reshape(-1, CHUNK)
        B = x.shape[0]
        outs = [np.zeros(s, np.float32) for s in ((B, 16, 25), (B, 16, 13), (B, 32, 7), (B, 32, 7), (B, 64, 7))]
        self._check(lib().silero_b200_stage_exact_pipeline(self._h, _p(x), B, *[_p(o) for o in outs]))
        return outs

    def stage_exact_layer(self, layer, x, want_y1=False):
        cin, c, t, stride = ((129, 16, 25, 2), (16, 32, 13, 2), (32, 32, 7, 1), (32, 64, 7, 1))[layer]
        x = _f32(x).reshape(-1, cin, t)
        B = x.shape[0]
        out = np.zeros((B, c, 1 + (t - 1) // stride), np.float32)
        y1 = np.zeros((B, 16, 25), np.float32) if (want_y1 and layer == 0) else None
        self._check(lib().silero_b200_stage_exact_layer(self._h, layer, _p(x), B, _p(out), _p(y1)))
        return (out, y1) if want_y1 else out

    def stage_exact_encoder(self, spec, kind=0):
        x = _f32(spec).reshape(-1, 129, 25)
        B = x.shape[0]
        outs = [np.zeros(s, np.float32) for s in ((B, 16, 13), (B, 32, 7), (B, 32, 7), (B, 64, 7))]
        self._check(lib().silero_b200_stage_exact_encoder(self._h, _p(x), B, kind, *[_p(o) for o in outs]))
        return outs

    def stage_exact_softmax(self, x):
        x = _f32(x)
        out = np.zeros_like(x)
        self._check(lib().silero_b200_stage_exact_softmax(self._h, _p(x), x.shape[0], x.shape[1], _p(out)))
        return out

    def stage_exact_lstm(self, x, h0=None, c0=None, wave=False):
        x = _f32(x).reshape(-1, 7, 64)
        B = x.shape[0]
        out, hn, cn = np.zeros((B, 7, 64), np.float32), np.zeros((2, 64), np.float32), np.zeros((2, 64), np.float32)
        h0 = _f32(h0).reshape(2, 64) if h0 is not None else None
        c0 = _f32(c0).reshape(2, 64) if c0 is not None else None
        self._check(lib().silero_b200_stage_exact_lstm(self._h, _p(x), B, _p(h0), _p(c0), _p(out), _p(hn), _p(cn), int(wave)))
        return out, hn, cn

    def stage_layer(self, layer, x):
        cin, c, t, stride = ((129, 16, 25, 2), (16, 32, 13, 2), (32, 32, 7, 1), (32, 64, 7, 1))[layer]
        x = _f32(x).reshape(-1, cin, t)
        out = np.zeros((x.shape[0], c, 1 + (t - 1) // stride), np.float32)
        self._check(lib().silero_b200_stage_layer(self._h, layer, _p(x), x.shape[0], _p(out)))
        return out

    def stage_layer_tap(self, layer, entry, tap, x):
        cin, c, t, stride = ((129, 16, 25, 2), (16, 32, 13, 2), (32, 32, 7, 1), (32, 64, 7, 1))[layer]
        x = _f32(x).reshape((-1, cin, t) if entry == 0 else (-1, t, c))
        out = np.zeros((x.shape[0], c, 1 + (t - 1) // stride) if tap == 0 else (x.shape[0], t, c), np.float32)
        self._check(lib().silero_b200_stage_layer_tap(self._h, layer, entry, tap, _p(x), x.shape[0], _p(out)))
        return out

    def stage_lstm(self, x, h0=None, c0=None):
        x = _f32(x).reshape(-1, 7, 64)
        out = np.zeros_like(x)
        hn = np.zeros((2, 64), np.float32)
        cn = np.zeros((2, 64), np.float32)
        h0 = _f32(h0).reshape(128) if h0 is not None else None
        c0 = _f32(c0).reshape(128) if c0 is not None else None
        self._check(lib().silero_b200_stage_lstm(self._h, _p(x), x.shape[0], _p(h0), _p(c0), _p(out), _p(hn), _p(cn)))
        return out, hn, cn

    def stage_decoder(self, x):
        x = _f32(x).reshape(-1, 64, 7)
        out = np.zeros((x.shape[0], 2), np.float32)
        self._check(lib().silero_b200_stage_decoder(self._h, _p(x), x.shape[0], _p(out)))
        return out


class Group:
    """Several GPUs in one host process (silero_b200_group_*, csrc/group.c): streams sharded in contiguous blocks, one host thread
    per device, results gathered in the caller's arrays."""

    def __init__(self, devices, max_streams, weights=None, **kw):
        L = lib()
        L.silero_b200_group_last_error.restype = C.c_char_p
        opts = Opts()
        L.silero_b200_default_opts(C.byref(opts))
        opts.max_streams = max_streams
        for k, v in kw.items():
            setattr(opts, k, v)
        devs = (C.c_int * len(devices))(*devices)
        self._g = C.c_void_p()
        if weights is None:
            weights = WEIGHTS_PATH
        if isinstance(weights, (bytes, bytearray)):
            rc = L.silero_b200_group_create(bytes(weights), C.c_size_t(len(weights)), devs, len(devices), C.byref(opts), C.byref(self._g))
        else:
            rc = L.silero_b200_group_create_from_file(os.fsencode(weights), devs, len(devices), C.byref(opts), C.byref(self._g))
        self._check(rc)
        self.max_streams = max_streams

    def _check(self, rc):
        if rc != 0:
            raise EngineError("silero_b200_group error %d: %s" % (rc, lib().silero_b200_group_last_error().decode()))

    def close(self):
        if self._g:
            lib().silero_b200_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n, per, mx = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().silero_b200_group_get_info(self._g, C.byref(n), C.byref(per), C.byref(mx)))
        return {"ndevices": n.value, "streams_per_device": per.value, "max_streams": mx.value}

    def segments_configure(self, params=None):
        self._check(lib().silero_b200_group_segments_configure(self._g, C.byref(params) if params is not None else None))

    def reset(self, first_stream=0, nstreams=None):
        self._check(lib().silero_b200_group_reset(self._g, first_stream, self.max_streams - first_stream if nstreams is None else nstreams))

    def run_streams_segments(self, pcm, nchunks=None, end_of_stream=False, cap=None, first_stream=0):
        """pcm: int16 [S, nsamples]. Returns (probs [S, nchunks], list of per-stream [(start_chunk, end_chunk), ...])."""
        assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.strides[1] == 2
        S = pcm.shape[0]
        if nchunks is None:
            nchunks = pcm.shape[1] // CHUNK
        if cap is None:
            cap = nchunks // 2 + 2
        probs = np.zeros((S, nchunks), np.float32)
        segs = np.zeros((S, cap, 2), np.int32)
        counts = np.zeros(S, np.int32)
        self._check(lib().silero_b200_group_run_streams_segments(self._g, _p(pcm), C.c_longlong(pcm.strides[0] // 2), first_stream, S, nchunks,
                                                                 int(end_of_stream), _p(segs), cap, _p(counts), _p(probs)))
        return probs, [[tuple(p) for p in segs[s, :min(counts[s], cap)].tolist()] for s in range(S)]

    def run_streams_segments_ptr(self, pcm_ptr, stream_stride, nstreams, nchunks, end_of_stream, segs_ptr, cap, counts_ptr, probs_ptr=None, first_stream=0):
        self._check(lib().silero_b200_group_run_streams_segments(self._g, C.c_void_p(pcm_ptr), C.c_longlong(stream_stride), first_stream, nstreams, nchunks,
                                                                 int(end_of_stream), C.c_void_p(segs_ptr), cap, C.c_void_p(counts_ptr),
                                                                 C.c_void_p(probs_ptr) if probs_ptr else None))


def pinned_empty(shape, dtype):
    """numpy view over cudaHostAlloc'ed memory (kept alive by the returned array's base object)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib().silero_b200_host_alloc_pinned(C.c_size_t(max(n, 1)), C.byref(p))
    if rc != 0:
        raise EngineError("pinned alloc failed: %s" % lib().silero_b200_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p.value


def pinned_free(ptr):
    lib().silero_b200_host_free_pinned(C.c_void_p(ptr))


# ---- segmenter / synth (host C) -----------------------------------------------------------------
def seg_params(**kw):
    p = SegParams()
    lib().vadc_seg_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def segments_text(prob, params=None):
    prob = _f32(prob).reshape(-1)
    params = params or seg_params()
    cap = 64 * (len(prob) + 2) + 64
    buf = C.create_string_buffer(cap)
    n = lib().vadc_segments_text(_p(prob), C.c_longlong(len(prob)), C.byref(params), buf, C.c_size_t(cap))
    return buf.raw[:n].decode()


class StreamSegmenter:
    """Streaming use of vadc_segmenter_* (feed in pieces, finish at end of stream)."""

    def __init__(self, params=None):
        self.s = Segmenter()
        self.params = params or seg_params()
        lib().vadc_segmenter_init(C.byref(self.s), C.byref(self.params))

    def _collect(self, fn, *args):
        cap = 4096
        arr = (Segment * cap)()
        n = fn(C.byref(self.s), *args, arr, C.c_longlong(cap))
        assert n <= cap
        return [(arr[i].start_chunk, arr[i].end_chunk) for i in range(n)]

    def feed(self, prob):
        prob = _f32(prob).reshape(-1)
        out = []
        for i in range(0, len(prob), 2048):
            piece = prob[i:i + 2048]
            out += self._collect(lib().vadc_segmenter_feed, _p(piece), C.c_longlong(len(piece)))
        return out

    def finish(self):
        return self._collect(lib().vadc_segmenter_finish)

    def format(self, seg):
        buf = C.create_string_buffer(96)
        n = lib().vadc_segment_format(C.byref(self.s), Segment(*seg), buf, C.c_size_t(96))
        return buf.raw[:n].decode()


_synth = None


def synth_pcm(seed, nsamples, kind=0):
    """Deterministic synthetic s16le test audio (csrc/synth.c). Host C in its own small library: generating input never loads the
    CUDA engine (bench.py's CPU reference arm depends on that)."""
    global _synth
    if _synth is None:
        path = os.path.join(_HERE, "libvadc_synth.so")
        if not os.path.exists(path):
            raise EngineError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        _synth = C.CDLL(path)
    out = np.zeros(nsamples, np.int16)
    _synth.vadc_synth_pcm(C.c_ulonglong(seed), kind, C.c_longlong(nsamples), _p(out))
    return out
