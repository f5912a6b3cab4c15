"""GPU: the on-device segmenter (vadc_b200/csrc/segment_kernel.cuh) against the host state machine
(segmenter.c, itself pinned byte-for-byte to the reference CLI's stdout in test_host_logic.py) and
against the golden stdout of the unmodified reference. Bar: (start_chunk, end_chunk) pairs bit-exact,
whatever the split of a stream into calls."""
import glob
import os

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))


def host_pairs(prob, params=None):
    s = vadc_b200.StreamSegmenter(params)
    return s.feed(prob) + s.finish()


def text_of(pairs, params=None):
    s = vadc_b200.StreamSegmenter(params)
    return "".join(s.format(p) for p in pairs)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_device_segments_reproduce_reference_stdout(engine, path):
    g = np.load(path)
    pcm = g["pcm"]
    engine.reset()
    engine.segments_configure()
    segs, counts, probs = engine.run_streams_segments(pcm[None, :], end_of_stream=True, want_probs=True)
    assert text_of(segs[0]) == str(g["stdout"])
    assert segs[0] == host_pairs(probs[0])
    # centiseconds is only a formatting option (vadc.c:251-256)
    assert text_of(segs[0], vadc_b200.seg_params(centiseconds=1)) == str(g["stdout_centi"])


def test_device_segments_streaming_many_streams(engine):
    """32 streams fed in uneven pieces; state (FeedState + buffered candidate) persists on the device."""
    S, N = 32, 260
    pcm = np.stack([vadc_b200.synth_pcm(700 + s, 1536 * N, kind=(2 if s == 5 else 1 if s == 6 else 0)) for s in range(S)])
    engine.reset()
    engine.segments_configure()
    whole, _, probs = engine.run_streams_segments(pcm, end_of_stream=True, want_probs=True)
    for s in range(S):
        assert whole[s] == host_pairs(probs[s]), s
    assert sum(len(w) for w in whole) > S          # the synthetic speech produces segments
    # same streams in pieces of 1, 7, 100, 152 chunks, then an empty end-of-stream call
    engine.reset()
    engine.segments_reset()
    got = [[] for _ in range(S)]
    n0 = 0
    for n in (1, 7, 100, 152):
        part, counts = engine.run_streams_segments(pcm[:, n0 * 1536:(n0 + n) * 1536])
        assert (counts <= n // 2 + 2).all()
        for s in range(S):
            got[s] += part[s]
        n0 += n
    assert n0 == N
    part, _ = engine.run_streams_segments(None, end_of_stream=True, first_stream=0)
    for s in range(S):
        got[s] += part[s]
        assert got[s] == whole[s], s


def test_device_segmenter_alone_matches_host_on_adversarial_probabilities(engine):
    """Random-walk probabilities hugging both thresholds, non-default options, capacity overflow."""
    import ctypes as C
    rng = np.random.default_rng(11)
    S, N = 64, 4000
    steps = rng.normal(0, 0.12, size=(S, N)).astype(np.float32)
    p = np.clip(0.42 + np.cumsum(steps, axis=1) * 0.2, 0, 1).astype(np.float32)
    p[:, ::97] = np.float32(0.5)                     # exactly at the threshold
    p[:, 5::89] = np.float32(0.5) - np.float32(0.15)  # exactly at the negative threshold
    p[3] = 0.9                                        # speech to the very end: closed by the end-of-stream rule
    p[4] = 0.1
    for params in (None, vadc_b200.seg_params(min_silence_ms=500.0, min_speech_ms=96.0, threshold=0.6, speech_pad_ms=200.0)):
        engine.segments_configure(params)
        cap = 1024
        d_p = engine.device_alloc(p.nbytes)
        d_s = engine.device_alloc(S * cap * 8)
        d_c = engine.device_alloc(S * 4)
        engine.h2d(d_p, p)
        half = N // 2
        segs = np.zeros((2, S, cap, 2), np.int32)
        cnt = np.zeros((2, S), np.int32)
        # two calls: first half, then second half with end of stream (row stride N, column offset via pointer)
        engine.segment_probs_device(d_p, N, S, half, False, d_s, cap, d_c)
        engine.sync()
        engine.d2h(segs[0], d_s); engine.d2h(cnt[0], d_c)
        engine.segment_probs_device(d_p + half * 4, N, S, N - half, True, d_s, cap, d_c)
        engine.sync()
        engine.d2h(segs[1], d_s); engine.d2h(cnt[1], d_c)
        total = 0
        for s in range(S):
            got = [tuple(int(v) for v in segs[k, s, i]) for k in range(2) for i in range(cnt[k, s])]
            assert got == host_pairs(p[s], params), s
            total += len(got)
        assert total > 100
        # overflow is counted, not stored
        engine.segments_configure(params)
        engine.segment_probs_device(d_p, N, S, N, True, d_s, 1, d_c)
        engine.sync()
        c1 = np.zeros(S, np.int32); engine.d2h(c1, d_c)
        assert [int(c) for c in c1] == [len(host_pairs(p[s], params)) for s in range(S)]
        for d in (d_p, d_s, d_c):
            engine.device_free(d)
    engine.segments_configure()


def test_async_submit_wait_equals_synchronous_calls():
    """silero_b200_submit_streams_segments / silero_b200_wait: six calls submitted back to back (more than the four
    tickets that may be in flight), each with its own pinned buffers; outputs identical to the blocking calls."""
    S, n, calls = 24, 40, 6
    pcm = np.stack([vadc_b200.synth_pcm(1300 + s, 1536 * n * calls) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, window_chunks=16)
    e.segments_configure()
    want_p, want_s = [], []
    for k in range(calls):
        s, c, p = e.run_streams_segments(np.ascontiguousarray(pcm[:, k * n * 1536:(k + 1) * n * 1536]), end_of_stream=(k == calls - 1), want_probs=True)
        want_p.append(p)
        want_s.append(s)
    e.reset()
    e.segments_reset()
    cap = n // 2 + 2
    bufs, tickets = [], []
    for k in range(calls):
        hin, hin_ptr = vadc_b200.pinned_empty((S, n * 1536), np.int16)
        hin[:] = pcm[:, k * n * 1536:(k + 1) * n * 1536]
        hp, hp_ptr = vadc_b200.pinned_empty((S, n), np.float32)
        hs, hs_ptr = vadc_b200.pinned_empty((S, cap, 2), np.int32)
        hc, hc_ptr = vadc_b200.pinned_empty((S,), np.int32)
        bufs.append((hin_ptr, hp, hp_ptr, hs, hs_ptr, hc, hc_ptr))
        tickets.append(e.submit_streams_segments_ptr(hin_ptr, n * 1536, S, n, k == calls - 1, hs_ptr, cap, hc_ptr, hp_ptr))
    for k in reversed(range(calls)):                       # waiting out of order is allowed
        e.wait(tickets[k])
    for k in range(calls):
        _, hp, _, hs, _, hc, _ = bufs[k]
        assert np.array_equal(hp, want_p[k]), k
        got = [[(int(a), int(b)) for a, b in hs[s, :hc[s]]] for s in range(S)]
        assert got == want_s[k], k
    with pytest.raises(vadc_b200.EngineError):
        e.wait(10 ** 6)
    for b in bufs:
        for ptr in (b[0], b[2], b[4], b[6]):
            vadc_b200.pinned_free(ptr)
    e.close()
