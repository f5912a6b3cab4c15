"""GPU: end-to-end parity of the CUDA path (through the C ABI) with the oracle and with the golden
vectors generated from the unmodified reference. Bars (BASELINE.json north_star): per-chunk
probabilities within 1e-4 max abs error, segment text bit-exact; STFT magnitudes bit-exact."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, Oracle, have_ref, ref_cli

pytestmark = pytest.mark.gpu
PTOL = 1e-4
HYB_REL = 3e-4   # hybrid STFT: bins above the fix-up threshold carry <= c*eps/k_rel relative error (measured 7e-5)
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))


def f32(pcm):
    n = len(pcm) // 1536
    return (pcm[: n * 1536].astype(np.float32) / np.float32(32768.0)).reshape(n, 1536)


def margins(p):
    return float(np.abs(p - 0.5).min()), float(np.abs(p - np.float32(0.5) + np.float32(0.15)).min())


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_golden_reference_vectors(engine, engine_exact, path):
    g = np.load(path)
    pcm = g["pcm"]
    engine.reset()
    probs, out2 = engine.run_streams(pcm[None, :], want_out2=True)
    assert out2.shape[1:] == g["out2"].shape            # trailing partial chunk dropped (vadc.c:964)
    err = np.abs(out2[0] - g["out2"]).max()
    assert err <= PTOL, err
    assert np.array_equal(probs[0], out2[0, :, 1])
    assert vadc_b200.segments_text(probs[0]) == str(g["stdout"])
    assert vadc_b200.segments_text(probs[0], vadc_b200.seg_params(centiseconds=1)) == str(g["stdout_centi"])
    # per-stage tensors of the first 8 chunks
    x = f32(pcm[: 8 * 1536])
    assert np.array_equal(engine_exact.stage_stft_magnitude(x), g["stages_stft"])    # exact mode: bit-exact
    norm, logmag = engine_exact.stage_stft_norm(x)
    assert np.abs(norm - g["stages_norm"]).max() < 5e-6
    mag = engine.stage_stft_magnitude(x)                                              # hybrid mode (default)
    assert (np.abs(mag - g["stages_stft"]) <= HYB_REL * g["stages_stft"]).all()
    norm, logmag = engine.stage_stft_norm(x)
    assert np.abs(norm - g["stages_norm"]).max() < 2 * HYB_REL
    engine_exact.reset()
    assert np.abs(engine_exact.run_streams(pcm[None, :], want_out2=True)[1][0] - g["out2"]).max() <= PTOL
    for got, k in zip(engine.stage_pipeline(x), ("l1", "l2", "l3", "l4")):
        ref = g["stages_" + k]
        assert np.abs(got - ref).max() <= 2e-4 * max(1.0, float(np.abs(ref).max())), k


def _edge_signals():
    rng = np.random.default_rng(0)
    x = np.zeros((8, 1536), np.float32)
    x[1] = 1.0 - 2.0 ** -15                      # full-scale DC
    x[2] = rng.uniform(-1, 1, 1536)              # white
    x[3, ::2] = 0.999; x[3, 1::2] = -1.0         # Nyquist
    x[4, 700] = 1.0                              # impulse
    x[5] = (rng.integers(-3, 4, 1536) / 32768.0) # near-silent LSB noise (worst case for log1p)
    x[6] = 0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(1536) / 16000.0)   # pure tone on a bin centre: almost every bin is "small"
    x[7] = np.round(3000 * np.sin(2 * np.pi * 440.0 * np.arange(1536) / 16000.0)) / 32768.0 + x[5]
    return x


def test_stft_bit_exact_on_edge_signals(engine_exact, oracle):
    x = _edge_signals()
    oracle.reset()
    st = oracle.run_stages(x)
    assert np.array_equal(engine_exact.stage_stft_magnitude(x), st["stft"])
    norm, _ = engine_exact.stage_stft_norm(x)
    assert np.abs(norm - st["norm"]).max() < 5e-6


def test_hybrid_stft_on_edge_signals(oracle):
    """Hybrid mode (opt-in fast family): bins below the threshold are bit-identical to the reference (exact tree), the
    others within HYB_REL; degenerate signals push most bins onto the exact path."""
    engine = vadc_b200.Engine(max_streams=64, stft_mode=vadc_b200.STFT_HYBRID)
    x = _edge_signals()
    oracle.reset()
    st = oracle.run_stages(x)
    engine.stft_stats(reset=True)
    mag = engine.stage_stft_magnitude(x)
    total, exact = engine.stft_stats(reset=True)
    assert total == 8 * 129 * 25 and exact > 0.3 * total
    assert (np.abs(mag - st["stft"]) <= HYB_REL * st["stft"]).all()
    norm, _ = engine.stage_stft_norm(x)
    assert np.abs(norm - st["norm"]).max() < 2 * HYB_REL
    engine.reset()
    assert np.abs(engine.run_chunks(x) - st["out"]).max() <= PTOL
    engine.close()


def test_hybrid_exact_path_alone_is_bit_exact(oracle):
    """k_rel = huge sends every bin through the warp-cooperative exact tree of the hybrid kernel."""
    e = vadc_b200.Engine(stft_mode=vadc_b200.STFT_HYBRID, stft_k_rel=1e30)
    x = np.concatenate([_edge_signals(), f32(vadc_b200.synth_pcm(8, 1536 * 8))])
    oracle.reset()
    assert np.array_equal(e.stage_stft_magnitude(x), oracle.run_stages(x)["stft"])
    total, exact = e.stft_stats()
    assert exact > 0.85 * total   # all-zero frames have threshold 0 and are already exact (FFT of zeros)
    e.close()


def test_hybrid_fix_fraction_on_speech_like_audio():
    pcm = np.stack([vadc_b200.synth_pcm(60 + s, 1536 * 100) for s in range(8)])
    engine = vadc_b200.Engine(max_streams=8, stft_mode=vadc_b200.STFT_HYBRID)   # (STFT_AUTO would take the exact kernel for 8 streams)
    engine.stft_stats(reset=True)
    engine.run_streams(pcm)
    total, exact = engine.stft_stats(reset=True)
    engine.close()
    assert total == 8 * 100 * 3225 and 0 < exact < 0.02 * total, (total, exact)


def test_small_batches_take_the_exact_stft(oracle):
    """STFT_AUTO: fewer streams than the tensor-core threshold run the exact STFT kernel -- magnitudes, log1p (the C library's bits,
    libm_exact.cuh) and the normalization scalar (the reference's summation order) are BIT-identical to the reference, so the whole
    normalized spectrogram is; the rest of the fp32 path stays within the 1e-4 bar on a 600-chunk stream."""
    pcm = vadc_b200.synth_pcm(50000 + 13 * 17, 1536 * 600)
    x = f32(pcm)
    oracle.reset()
    st = oracle.run_stages(x)
    e = vadc_b200.Engine(max_streams=2)
    e.stft_stats(reset=True)
    p, out2 = e.run_streams(pcm[None, :], want_out2=True)
    assert e.stft_stats()[0] == 0                              # no bin went through the FFT kernels
    e.close()
    assert np.abs(out2[0] - st["out"]).max() <= PTOL
    assert vadc_b200.segments_text(p[0]) == oracle.segments_text(st["out"][:, 1])
    ex = vadc_b200.Engine(stft_mode=vadc_b200.STFT_EXACT)
    norm, logmag = ex.stage_stft_norm(x)
    ex.close()
    assert np.array_equal(norm.view(np.uint32), st["norm"].view(np.uint32))


@pytest.mark.parametrize("S,N,window", [(1, 1, 0), (3, 7, 0), (37, 70, 16), (130, 33, 5), (64, 96, 0)])
def test_multi_stream_vs_oracle(oracle, S, N, window):
    e = vadc_b200.Engine(max_streams=S, window_chunks=window)
    pcm = np.stack([vadc_b200.synth_pcm(1000 + s, N * 1536, kind=(0 if s % 11 else (1 if s % 2 else 2))) for s in range(S)])
    probs, out2 = e.run_streams(pcm, want_out2=True)
    worst = 0.0
    for s in sorted(set([0, 1, S // 2, S - 1]) | set(range(0, S, 13))):
        if s >= S:
            continue
        oracle.reset()
        ref = oracle.run_pcm(pcm[s])
        worst = max(worst, float(np.abs(out2[s] - ref).max()))
        assert vadc_b200.segments_text(probs[s]) == oracle.segments_text(ref[:, 1]), s
    assert worst <= PTOL, worst
    e.close()


def test_state_carries_across_calls_and_windows():
    S, N = 9, 50
    pcm = np.stack([vadc_b200.synth_pcm(50 + s, N * 1536) for s in range(S)])
    e1 = vadc_b200.Engine(max_streams=S, window_chunks=64)
    whole = e1.run_streams(pcm)
    e2 = vadc_b200.Engine(max_streams=S, window_chunks=7)      # ragged windows: 7*7+1
    assert np.array_equal(e2.run_streams(pcm), whole)
    e2.reset()
    a = e2.run_streams(np.ascontiguousarray(pcm[:, : 20 * 1536]))
    b = e2.run_streams(np.ascontiguousarray(pcm[:, 20 * 1536:]))
    assert np.array_equal(np.concatenate([a, b], 1), whole)    # split calls == one call, bit for bit
    # reset really zeroes, set/get round-trips
    h, c = e2.get_state(3)
    assert np.abs(h).max() > 0
    e2.reset()
    h0, c0 = e2.get_state(3)
    assert not h0.any() and not c0.any()
    e2.set_state(h, c, stream=3)
    h1, c1 = e2.get_state(3)
    assert np.array_equal(h, h1) and np.array_equal(c, c1)
    e1.close(); e2.close()


def test_run_chunks_is_backend_run(engine, oracle):
    """silero_b200_run_chunks == backend_run semantics: consecutive chunks of one stream, any batch,
    state carried; identical bits for batch 1 / 7 / 96 (finding F6) and for the s16 entry point."""
    pcm = vadc_b200.synth_pcm(77, 96 * 2 * 1536)
    x = f32(pcm)
    engine.reset()
    ref_all = engine.run_chunks(x, stream=5)
    for batch in (1, 7, 96):
        engine.reset()
        got = np.concatenate([engine.run_chunks(x[i:i + batch], stream=5) for i in range(0, len(x), batch)])
        assert np.array_equal(got, ref_all), batch
    engine.reset()
    probs, out2 = engine.run_streams(pcm[None, :], first_stream=5, want_out2=True)
    assert np.array_equal(out2[0], ref_all)
    oracle.reset()
    assert np.abs(ref_all - oracle.run_pcm(pcm)).max() <= PTOL
    # other streams' state untouched
    h, c = engine.get_state(4)
    assert not h.any() and not c.any()


def test_streams_are_independent_and_order_invariant():
    S, N = 40, 24
    base = [vadc_b200.synth_pcm(300 + i, N * 1536) for i in range(5)]
    pcm = np.stack([base[s % 5] for s in range(S)])
    e = vadc_b200.Engine(max_streams=S)
    p = e.run_streams(pcm)
    for s in range(S):
        assert np.array_equal(p[s], p[s % 5])          # duplicates in different slots / tiles agree bit for bit
    perm = np.random.default_rng(1).permutation(S)
    e.reset()
    assert np.array_equal(e.run_streams(np.ascontiguousarray(pcm[perm])), p[perm])
    e.close()


def test_device_resident_path_equals_host_path():
    S, N = 33, 40
    pcm = np.stack([vadc_b200.synth_pcm(900 + s, N * 1536) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, window_chunks=16)
    host = e.run_streams(pcm)
    e.reset()
    d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm)
    e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
    e.sync()
    dev = np.zeros((S, N), np.float32)
    e.d2h(dev, d_probs)
    assert np.array_equal(dev, host)
    ms, launches = e.last_timing()
    # windows of 16 + 16 + 8 chunks x 33 streams: 528 chunks take the thread-per-token encoder (STFT, front, 4 layers, LSTM wavefront,
    # decoder head = 8 launches), 264 chunks the CTA-per-chunk encoder (4 launches)
    assert launches == 8 + 8 + 4 and ms["total"] > 0
    e.device_free(d_pcm); e.device_free(d_probs); e.close()


def test_argument_errors(engine):
    L = vadc_b200.lib()
    with pytest.raises(vadc_b200.EngineError):
        engine.run_streams(np.zeros((65, 1536), np.int16))                 # more streams than max_streams=64
    with pytest.raises(vadc_b200.EngineError):
        engine.run_chunks(np.zeros((1, 1536), np.float32), stream=64)
    assert engine.run_streams(np.zeros((2, 100), np.int16)).shape == (2, 0)  # shorter than one chunk: nothing to do
    assert L.silero_b200_run_chunks(None, 0, None, 1, None) == -1
    i = engine.info()
    assert (i["batch_size_restriction"], i["is_silero_v5"], i["input_size_min"], i["input_size_max"], i["output_dims"]) == (-1, 0, 1536, 1536, 3)


def test_full_width_properties_and_sampled_parity(oracle):
    """BASELINE cfg3 width (4096 concurrent streams), shortened in time: size-independent properties
    (duplicate streams agree bit for bit across tiles, split == whole) plus oracle parity on a sample."""
    S, N, nb = 4096, 24, 16
    base = [vadc_b200.synth_pcm(7000 + i, N * 1536) for i in range(nb)]
    shift = lambda s: ((s // nb) * 5) % N
    pcm = np.stack([np.roll(base[s % nb].reshape(N, 1536), shift(s), 0).reshape(-1) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S)
    p = e.run_streams(pcm)
    groups = {}
    for s in range(S):
        groups.setdefault((s % nb, shift(s)), []).append(s)
    for members in groups.values():
        for s in members[1:]:
            assert np.array_equal(p[s], p[members[0]])
    e.reset()
    a = e.run_streams(np.ascontiguousarray(pcm[:, : 10 * 1536]))
    b = e.run_streams(np.ascontiguousarray(pcm[:, 10 * 1536:]))
    assert np.array_equal(np.concatenate([a, b], 1), p)
    worst = 0.0
    for s in (0, 17, 2049, 4095):
        oracle.reset()
        worst = max(worst, float(np.abs(p[s] - oracle.run_pcm(pcm[s])[:, 1]).max()))
    assert worst <= PTOL, worst
    e.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_timestamps_bit_exact_vs_reference_cli(engine):
    """cfg1: one 60 s synthetic stream: stdout of the unmodified reference CLI == our segments."""
    pcm = vadc_b200.synth_pcm(2024, 16000 * 60)
    engine.reset()
    p = engine.run_streams(pcm[None, :])[0]
    assert vadc_b200.segments_text(p) == ref_cli(pcm)
    raw = np.array([float(v) for v in ref_cli(pcm, "--raw_probabilities").split()])
    assert np.abs(raw - p).max() <= PTOL + 1e-6
    m = margins(p)
    print("threshold margins: min|p-0.5|=%.2e min|p-0.35|=%.2e" % m)


def test_cfg4_single_long_stream(oracle):
    """BASELINE cfg4, shortened in time: ONE stream, thousands of chunks. The stateless front end runs
    chunk-parallel over each window (tensor-core layers: >= 2048 chunks per pass), the LSTM is one serial
    scan with its state carried across windows (warm-up-overlap segmentation cannot be exact,
    so the engine does not segment the stream). Probabilities within 1e-4, timestamps identical."""
    N = 4500                                   # 7.2 minutes
    pcm = vadc_b200.synth_pcm(31337, N * 1536)
    e = vadc_b200.Engine(max_streams=1, window_chunks=2100)   # 3 windows: 2100 + 2100 + 300
    p, out2 = e.run_streams(pcm[None, :], want_out2=True)
    oracle.reset()
    ref = oracle.run_pcm(pcm)
    assert np.abs(out2[0] - ref).max() <= PTOL
    assert vadc_b200.segments_text(p[0]) == oracle.segments_text(ref[:, 1])
    # the same stream through the device segmenter, fed in two calls
    e.reset()
    e.segments_configure()
    s1, _ = e.run_streams_segments(pcm[None, : 3000 * 1536])
    s2, _ = e.run_streams_segments(pcm[None, 3000 * 1536:], end_of_stream=True)
    seg = vadc_b200.StreamSegmenter()
    assert "".join(seg.format(x) for x in s1[0] + s2[0]) == oracle.segments_text(ref[:, 1])
    e.close()


def test_cfg5_many_streams_sharded_like_ranks(oracle):
    """BASELINE cfg5, shortened in time: 16384 streams. Two engines each own one block of streams (what two
    ranks do, vadc_b200/shard.py); together they must reproduce one engine owning all streams bit for bit
    (no cross-stream coupling, tile composition is irrelevant), and the gathered segments must agree."""
    from vadc_b200 import shard
    S, N, nb = 16384, 6, 8
    base = [vadc_b200.synth_pcm(8100 + i, N * 1536) for i in range(nb)]
    pcm = np.stack([np.roll(base[s % nb].reshape(N, 1536), (s // nb) % N, 0).reshape(-1) for s in range(S)])
    whole = vadc_b200.Engine(max_streams=S)
    p_all = whole.run_streams(pcm)
    whole.close()
    parts, segs = [], []
    for rank in range(2):
        first, count = shard.stream_range(S, 2, rank)
        e = vadc_b200.Engine(max_streams=count)
        e.segments_configure()
        s, _, p = e.run_streams_segments(pcm[first:first + count], end_of_stream=True, want_probs=True)
        parts.append(p)
        segs += s
        e.close()
    assert np.array_equal(np.concatenate(parts, 0), p_all)
    for s in (0, 8191, 8192, 16383):
        oracle.reset()
        ref = oracle.run_pcm(pcm[s])[:, 1]
        assert np.abs(p_all[s] - ref).max() <= PTOL
        seg = vadc_b200.StreamSegmenter()
        assert "".join(seg.format(x) for x in segs[s]) == oracle.segments_text(ref)


def test_overlapped_and_single_stream_schedules_agree():
    """run_window overlaps the STFT of window w+1 (second CUDA stream) with the back end of window w; per-stage profiling keeps
    everything on one stream. Same bits either way, also across calls (the hand-off events persist) and for the host path, whose
    staging buffers are released by the STFT."""
    S, N = 96, 23
    pcm = np.stack([vadc_b200.synth_pcm(7000 + s, N * 1536) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, window_chunks=4)
    outs = []
    for prof in (0, 1, 0):
        e.set_profiling(prof)
        e.reset()
        a = e.run_streams(pcm[:, : 10 * 1536])               # two calls: state and events carried across
        b = e.run_streams(pcm[:, 10 * 1536:])
        outs.append(np.concatenate([a, b], axis=1))
    e.set_profiling(0)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    d_pcm = e.device_alloc(pcm.nbytes)
    d_probs = e.device_alloc(S * N * 4)
    e.h2d(d_pcm, pcm)
    e.reset()
    e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
    e.sync()
    got = np.zeros((S, N), np.float32)
    e.d2h(got, d_probs)
    assert np.array_equal(got, outs[0])
    e.close()


def test_lstm_nonlinearities_have_the_c_librarys_bits():
    """csrc/libm_exact.cuh: expf / tanhf of the fp32 LSTM path and log1pf of the exact STFT path equal the host C library's (the reference build's) bit for bit --
    random arguments over the gates' range, dense samples near zero and the saturation ends, special values."""
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = libm.tanhf.restype = libm.log1pf.restype = C.c_float
    libm.expf.argtypes = libm.tanhf.argtypes = libm.log1pf.argtypes = [C.c_float]
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-30, 30, 200000), rng.normal(0, 1, 100000), rng.uniform(-1e-3, 1e-3, 20000), rng.uniform(-87, 87, 50000),
                        np.float32([0.0, -0.0, 1.0, -1.0, 0.5, 22.0, -22.0, 21.999, 1e-20, -1e-20, 0.34657359, 1.0397208, 18.7, -18.7, 43.9])]).astype(np.float32)
    e = vadc_b200.Engine()
    ge, gt, _ = e.stage_libm(x)
    # log1pf over the range of magnitude * 2^20: denormal-small to 3e8, log-uniform, plus the branch points of the algorithm
    xl = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(3e8), 300000)), rng.uniform(0, 1, 50000), np.float32([0.0, 0.41421, 0.41422, 0.41423, 1.0, 2.0 ** -29, 2.0 ** 25, 3.4e8])]).astype(np.float32)
    gl = e.stage_libm(xl)[2]
    e.close()
    hl = np.array([libm.log1pf(float(v)) for v in xl], np.float32)
    assert np.array_equal(gl.view(np.uint32), hl.view(np.uint32))
    he = np.array([libm.expf(float(v)) for v in x], np.float32)
    ht = np.array([libm.tanhf(float(v)) for v in x], np.float32)
    assert np.array_equal(gt.view(np.uint32), ht.view(np.uint32))
    bad = np.nonzero(ge.view(np.uint32) != he.view(np.uint32))[0]
    assert bad.size == 0, (x[bad][:5], ge[bad][:5], he[bad][:5])
