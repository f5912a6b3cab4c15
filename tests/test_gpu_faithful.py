"""GPU: the faithful path (vadc_b200/csrc/faithful_kernel.cuh): stream batches of at most
SILERO_B200_FAITHFUL_MAX_STREAMS streams on a fully automatic engine -- the way the reference itself is used -- and any batch on
request. Bar: BIT-identical probabilities (both decoder outputs) and LSTM state against the oracle, which is itself pinned bit for
bit to the unmodified reference build (tests/test_oracle_vs_ref.py); hence identical timestamps for streams of any length."""
import glob
import os

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, Oracle, have_ref, ref_cli

CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "e2e_*.npz")))

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_out2(pcm):
    return Oracle().run_pcm(pcm)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_golden_vectors_of_the_unmodified_reference_bit_for_bit(path):
    """tests/golden/e2e_*.npz hold outputs of the UNMODIFIED reference build (tests/golden/make_e2e_golden.py): both decoder outputs
    of every chunk must come back with the same bits, and the segment text with the same bytes."""
    g = np.load(path)
    e = vadc_b200.Engine()
    p, out2 = e.run_streams(g["pcm"][None, :], want_out2=True)
    e.close()
    assert np.array_equal(bits(out2[0]), bits(g["out2"]))
    assert vadc_b200.segments_text(p[0]) == str(g["stdout"])


def test_cfg1_single_stream_60s_is_bit_identical():
    """BASELINE configs[0]: one 60 s stream (625 chunks) through run_streams, default engine."""
    pcm = vadc_b200.synth_pcm(4242, 1536 * 625)
    e = vadc_b200.Engine()
    p, out2 = e.run_streams(pcm[None, :], want_out2=True)
    h, c = e.get_state(0)
    e.close()
    o = Oracle()
    ref = o.run_pcm(pcm)
    assert np.array_equal(bits(out2[0]), bits(ref))
    assert np.array_equal(bits(p[0]), bits(ref[:, 1]))
    assert np.array_equal(bits(h).reshape(-1), bits(o.state[:128])) and np.array_equal(bits(c).reshape(-1), bits(o.state[128:]))


@pytest.mark.parametrize("S,N,window", [(1, 1, 0), (3, 7, 0), (8, 90, 17), (64, 25, 4), (128, 9, 0)])
def test_stream_batches_windows_and_partial_tiles(S, N, window):
    pcm = np.stack([vadc_b200.synth_pcm(600 + 7 * s, N * 1536) for s in range(S)])
    pcm[S // 2] = 0                                                        # an all-zero stream among them
    e = vadc_b200.Engine(max_streams=S, window_chunks=window)
    _, out2 = e.run_streams(pcm, want_out2=True)
    e.close()
    for s in range(S):
        assert np.array_equal(bits(out2[s]), bits(oracle_out2(pcm[s]))), s


def test_long_streams_stay_bit_identical():
    """The case the fast kernels cannot promise: thousands of chunks with long silences, where one ulp in
    the LSTM's forget gate is integrated into the cell state. 3000 chunks (4.8 min) per stream, several windows, two calls."""
    N = 3000
    pcm = np.stack([vadc_b200.synth_pcm(50000 + 13 * i, N * 1536) for i in (3, 11, 17)])
    e = vadc_b200.Engine(max_streams=3, window_chunks=700)
    a = e.run_streams(pcm[:, : 1100 * 1536], want_out2=True)[1]
    b = e.run_streams(pcm[:, 1100 * 1536:], want_out2=True)[1]
    e.close()
    out2 = np.concatenate([a, b], axis=1)
    for s in range(3):
        ref = oracle_out2(pcm[s])
        assert np.array_equal(bits(out2[s]), bits(ref)), (s, float(np.abs(out2[s] - ref).max()))


def test_run_chunks_is_bit_identical_to_backend_run():
    """silero_b200_run_chunks == backend_run (silero.h:53-74): batches of 96 chunks and a short last one, state carried."""
    pcm = vadc_b200.synth_pcm(99, 1536 * 230)
    x = (pcm.astype(np.float32) / np.float32(32768.0)).reshape(-1, 1536)
    e = vadc_b200.Engine()
    got = np.concatenate([e.run_chunks(x[i:i + 96]) for i in range(0, len(x), 96)])
    e.close()
    o = Oracle()
    want = np.concatenate([o.run_chunks(x[i:i + 96]) for i in range(0, len(x), 96)])
    assert np.array_equal(bits(got), bits(want))


def test_default_for_any_batch_and_not_taken_by_an_explicit_fast_family():
    """The kernel family is a property of the engine, fixed at creation: a default engine runs the exact path for ANY number of streams
    (bit-identical whatever the batch composition); an engine created with an explicit fast mode never does (1e-4 bar)."""
    S, N = 130, 12
    pcm = np.stack([vadc_b200.synth_pcm(3000 + s, N * 1536) for s in range(S)])
    ref = np.stack([oracle_out2(pcm[s]) for s in range(S)])
    for kw in ({}, {"layer_mode": vadc_b200.LAYERS_FAITHFUL}):
        e = vadc_b200.Engine(max_streams=S, **kw)
        out2 = e.run_streams(pcm, want_out2=True)[1]
        # the same streams again in other batch shapes: 4 at a time, then one by one -- same engine, same bits
        e.reset()
        few = np.concatenate([e.run_streams(pcm[i:i + 4], want_out2=True, first_stream=i)[1] for i in range(0, 8, 4)])
        e.reset()
        one = e.run_streams(pcm[5:6], want_out2=True, first_stream=5)[1]
        e.close()
        assert np.array_equal(bits(out2), bits(ref))
        assert np.array_equal(bits(few), bits(ref[:8])) and np.array_equal(bits(one), bits(ref[5:6]))
    e = vadc_b200.Engine(max_streams=S, lstm_mode=vadc_b200.LSTM_FP32)
    fast = e.run_streams(pcm[:4], want_out2=True)[1]
    launches = e.last_timing()[1]
    e.close()
    assert launches % 7 == 0 and np.abs(fast - ref[:4]).max() <= 1e-4            # 7 kernels per window: the fast family
    assert not np.array_equal(bits(fast), bits(ref[:4]))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_raw_probabilities_text_equals_the_reference_cli():
    """The unmodified reference CLI's --raw_probabilities output ("%f" per chunk, vadc.c:992-997) from the engine's numbers."""
    pcm = vadc_b200.synth_pcm(2024, 1536 * 400)
    e = vadc_b200.Engine()
    p = e.run_streams(pcm[None, :])[0]
    e.close()
    assert "".join("%f\n" % v for v in p) == ref_cli(pcm, "--raw_probabilities")


def test_a_lost_wavefront_producer_is_an_error_not_a_wrong_answer():
    """exact_lstm_kernel's two-layer wavefront (few streams): the layer-1 CTA of a stream group polls its layer-0 partner's progress. With the test hook the producer
    never publishes: the consumer gives up after the poll limit, the CALL FAILS (SILERO_B200_ERR_CUDA, no silent garbage) and the
    stream's persistent state is left as it was; the engine is usable again afterwards."""
    pcm = vadc_b200.synth_pcm(99, 30 * 1536)[None, :]
    e = vadc_b200.Engine(max_streams=1)
    good = e.run_streams(pcm, want_out2=True)[1]
    h0, c0 = e.get_state(0)
    e.debug_wavefront(stall_producer=True, spin_limit=2000)
    with pytest.raises(vadc_b200.EngineError, match="lost its layer-0 producer"):
        e.run_streams(pcm)
    h1, c1 = e.get_state(0)
    assert np.array_equal(bits(h1[1]), bits(h0[1])) and np.array_equal(bits(c1[1]), bits(c0[1]))   # layer 1 (the consumer) kept its state
    e.debug_wavefront()
    e.reset()
    again = e.run_streams(pcm, want_out2=True)[1]
    e.close()
    assert np.array_equal(bits(again), bits(good))
