"""GPU: the exact path at scale -- the kernels a default engine runs for ANY number of streams (stft_sym_kernel, exact_front_kernel,
exact_layer_kernel, exact_lstm_kernel). Everything here is compared bit for bit: with the oracle (pinned to the unmodified reference
build, tests/test_oracle_vs_ref.py), with the reference's own golden fixtures (1e-4, the reference's atol), and between the kernel
mappings that serve different batch shapes."""
import multiprocessing as mp
import os

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, Oracle
from testtensor_io import load_list
from test_gpu_fixtures import ATOL, blob, fx

pytestmark = pytest.mark.gpu
CHUNK = 1536


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def f32(pcm):
    return (pcm.astype(np.float32) / np.float32(32768.0)).reshape(-1, CHUNK)


def _oracle_job(pcm):
    return Oracle().run_pcm(pcm)


def oracle_many(streams):
    """The oracle on several streams, one process per core (it is a scalar C port: ~235 x realtime per core)."""
    n = min(len(streams), len(os.sched_getaffinity(0)), 16)
    with mp.get_context("spawn").Pool(n) as pool:
        return pool.map(_oracle_job, streams)


# ---- stage by stage ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 23, 1100])
def test_encoder_stages_bit_identical_to_the_oracle(B):
    """exact STFT -> front -> four layers, every layer output: 0 differing bits (1100 chunks: several batches per persistent CTA)."""
    x = f32(vadc_b200.synth_pcm(4242 + B, CHUNK * B))
    ref = Oracle().run_stages(x)
    e = vadc_b200.Engine()
    y1, l1, l2, l3, l4 = e.stage_exact_pipeline(x)
    e.close()
    for name, got in (("l1", l1), ("l2", l2), ("l3", l3), ("l4", l4)):
        assert np.array_equal(bits(got), bits(ref[name])), name


def test_layer_taps_from_the_oracles_own_stage_tensors():
    """Each layer alone, fed with the oracle's input of that layer: isolates a layer from its predecessors."""
    x = f32(vadc_b200.synth_pcm(77, CHUNK * 40))
    ref = Oracle().run_stages(x)
    e = vadc_b200.Engine()
    for layer, (src, dst) in enumerate((("norm", "l1"), ("l1", "l2"), ("l2", "l3"), ("l3", "l4"))):
        assert np.array_equal(bits(e.stage_exact_layer(layer, ref[src])), bits(ref[dst])), layer
    for kind, src in ((1, "norm"),):
        outs = e.stage_exact_encoder(ref[src], kind=kind)
        assert np.array_equal(bits(outs[3]), bits(ref["l4"]))
    e.close()


def test_mirrored_stft_equals_the_full_tree_kernel_and_the_oracle(monkeypatch):
    """stft_sym_kernel (128 shared rows, plain + alternating lane sums) against stft_logmag_kernel (all 256 rows) and the oracle on
    edge signals + speech: identical magnitudes, bit for bit."""
    from test_gpu_parity import _edge_signals
    x = np.concatenate([_edge_signals(), f32(vadc_b200.synth_pcm(8, CHUNK * 40))])
    ref = Oracle().run_stages(x)["stft"]
    e = vadc_b200.Engine()
    sym = e.stage_stft_magnitude(x)
    e.close()
    monkeypatch.setenv("SILERO_B200_STFT_NO_SYM", "1")
    e = vadc_b200.Engine()
    full = e.stage_stft_magnitude(x)
    e.close()
    assert np.array_equal(bits(sym), bits(ref)) and np.array_equal(bits(full), bits(ref))


def test_a_basis_without_the_mirror_property_takes_the_full_tree():
    """The engine checks the table when it is created: one perturbed entry and the exact STFT runs on all 256 rows (still exact)."""
    from testtensor_io import load_testtensor
    basis = np.array(list(load_testtensor(vadc_b200.WEIGHTS_PATH).values())[0], np.float32).reshape(258, 256).copy()
    basis[37, 91] *= np.float32(1.0 + 2.0 ** -20)
    w = blob({0: basis})
    x = f32(vadc_b200.synth_pcm(5, CHUNK * 6))
    e = vadc_b200.Engine(weights=w)
    got = e.stage_stft_magnitude(x)
    e.close()
    ref = Oracle(weights=w).run_stages(x)["stft"]
    assert np.array_equal(bits(got), bits(ref))


# ---- the reference's golden fixtures on the kernels that serve traffic -------------------------------------------------------
def engine(overrides):
    return vadc_b200.Engine(weights=blob(overrides), max_streams=1)


def test_fixture_transformer_first_layer():
    v = fx("transformer_first_layer")
    e = engine({1 + i: v[i] for i in range(24)})
    assert np.abs(e.stage_exact_layer(0, v[24]) - v[25]).max() < ATOL


def test_fixture_transformer_layers_3():
    v = fx("transformer_layers_3")
    e = engine({49 + i: v[i] for i in range(22)})
    assert np.abs(e.stage_exact_layer(2, v[22]) - v[23]).max() < ATOL


@pytest.mark.parametrize("name,nl", [("transformer_layers_1_2", 2), ("transformer_layers_1_2_3", 3), ("transformer_layers_1_2_3_4", 4)])
def test_fixture_cumulative_layers(name, nl):
    v = fx(name)
    nw = (24, 48, 70, 94)[nl - 1]
    e = engine({1 + i: v[i] for i in range(nw)})
    assert np.abs(e.stage_exact_encoder(v[nw], kind=1)[nl - 1] - v[nw + 1]).max() < ATOL


def test_fixture_adaptive_normalization_encoder():
    v = fx("adaptive_normalization_encoder")
    e = engine({1 + i: v[i] for i in range(94)})
    assert np.abs(e.stage_exact_encoder(v[94], kind=2)[3] - v[95]).max() < ATOL


def test_fixture_first_layer_conv_block():
    """64-frame fixture through the 25-frame front kernel as three overlapping windows (see test_gpu_fixtures.conv_block_t64)."""
    dw_w, dw_b, pw_w, pw_b, pr_w, pr_b, x, exp = fx("first_layer_conv_block")
    e = engine({1: dw_w, 2: dw_b, 3: pw_w, 4: pw_b, 5: pr_w, 6: pr_b})
    batch = np.stack([x[:, s:s + 25] for s in (0, 21, 39)])
    y = e.stage_exact_layer(0, batch, want_y1=True)[1]               # [3,16,25]
    out = np.zeros((16, 64), np.float32)
    out[:, 0:23], out[:, 23:44], out[:, 41:64] = y[0][:, 0:23], y[1][:, 2:23], y[2][:, 2:25]
    assert np.abs(out - exp).max() < ATOL


def test_fixture_softmax():
    """The reference's 100 x 100 softmax fixture (test.c:900) on the GPU: the attention's softmax arithmetic (tensor.h:751-784) as a
    row tap; also bit-identical to the oracle's restatement."""
    import ctypes as C
    x, exp = fx("softmax_test")
    e = vadc_b200.Engine()
    got = e.stage_exact_softmax(np.asarray(x, np.float32).reshape(100, 100))
    e.close()
    assert np.abs(got - np.asarray(exp).reshape(100, 100)).max() < ATOL
    o = Oracle()
    y = np.ascontiguousarray(np.asarray(x, np.float32).reshape(100, 100)).copy()
    o.lib.so_softmax_rows(y.ctypes.data_as(C.c_void_p), 100, 100)
    assert np.array_equal(bits(got), bits(y))


def test_expf_underflow_range():
    """expf between the last normal result and the underflow to zero (-87.3 .. -103.97): glibc rounds the double result to a denormal,
    takes __math_may_uflowf below log(2^-149) and returns 0 below log(2^-150). Round 1's restatement returned 0 from -103 on (found by
    the softmax fixture, whose rows span +-450)."""
    import ctypes as C
    lib = C.CDLL("libm.so.6")
    lib.expf.restype, lib.expf.argtypes = C.c_float, [C.c_float]
    x = np.concatenate([np.linspace(-104.5, -86.0, 20001).astype(np.float32), np.float32([-103.97208, -103.972076, -103.2789, -103.27893, 88.7, 88.73])])
    want = np.array([lib.expf(float(v)) for v in x], np.float32)
    e = vadc_b200.Engine()
    got = e.stage_libm(x)[0]
    e.close()
    assert np.array_equal(bits(got), bits(want))


def test_fixture_lstm():
    x, h0, c0, w, b, exp = fx("lstm_nito_reference_randn")
    e = engine({95: w, 96: b})
    out, hn, cn = e.stage_exact_lstm(x.reshape(1, 7, 64), h0, c0)
    assert np.abs(np.concatenate([out.reshape(7, 64), hn, cn], 0) - exp).max() < ATOL


def test_lstm_kernels_agree_bit_for_bit():
    """One stream: the multi-stream kernel (weights in registers) and the two-layer wavefront launch of the same kernel on a long
    sequence from a non-zero state: same outputs, same final state."""
    rng = np.random.default_rng(3)
    x = np.maximum(rng.standard_normal((60, 7, 64)).astype(np.float32), 0)
    h0, c0 = rng.standard_normal((2, 64)).astype(np.float32) * 0.3, rng.standard_normal((2, 64)).astype(np.float32)
    e = vadc_b200.Engine()
    a = e.stage_exact_lstm(x, h0, c0)
    b = e.stage_exact_lstm(x, h0, c0, wave=True)
    e.close()
    for u, v in zip(a, b):
        assert np.array_equal(bits(u), bits(v))


# ---- the whole path ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,N,window", [(75, 20, 7), (75, 20, 5), (12, 31, 0), (12, 32, 0), (149, 12, 0), (149, 10, 0), (300, 30, 11), (1185, 9, 4)])
def test_batch_shapes_around_the_kernel_switches(S, N, window):
    """Stream counts around the points where the engine changes mapping (LSTM wavefront launch <-> one launch per layer at 10 x SMs / 2 streams,
    CTA-per-chunk <-> thread-per-token encoder at 384 chunks per window -- 375 / 372 chunks below it, 384 / 525 above; layer batches of
    fewer chunks than a CTA holds on small windows; one and two streams per LSTM CTA): identical bits."""
    base = [vadc_b200.synth_pcm(900 + 17 * i, CHUNK * N) for i in range(16)]
    pcm = np.stack([np.roll(base[s % 16], CHUNK * ((s // 16) % N)) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, window_chunks=window)
    out2 = e.run_streams(pcm, want_out2=True)[1]
    e.close()
    sample = sorted(set(list(range(0, S, max(1, S // 7))) + [S - 1]))
    for s, ref in zip(sample, oracle_many([pcm[s] for s in sample])):
        assert np.array_equal(bits(out2[s]), bits(ref)), s


def test_long_streams_at_scale_gate():
    """THE parity gate of the benchmarked configuration: 1024 streams x 3000 chunks (4.8 minutes each, long silences included) on the
    DEFAULT engine, fed in 24 calls of 125 chunks with the state carried on the device; 32 sampled streams against the oracle:
    probabilities bit-identical (bar: 1e-4), segment text identical."""
    S, N, STEP, NB = 1024, 3000, 125, 32
    base = np.stack([vadc_b200.synth_pcm(50000 + 13 * i, N * CHUNK) for i in range(NB)]).reshape(NB, N, CHUNK)
    sid, off = np.arange(S) % NB, (np.arange(S) // NB) * 37
    e = vadc_b200.Engine(max_streams=S)
    probs = np.zeros((S, N), np.float32)
    for k in range(N // STEP):
        idx = (off[:, None] + k * STEP + np.arange(STEP)[None, :]) % N
        pcm = base[sid[:, None], idx].reshape(S, STEP * CHUNK)
        probs[:, k * STEP:(k + 1) * STEP] = e.run_streams(pcm)
    e.close()
    sample = [33 * j for j in range(31)] + [S - 1]
    streams = [base[sid[s]][(off[s] + np.arange(N)) % N].reshape(-1) for s in sample]
    worst = 0.0
    for s, ref in zip(sample, oracle_many(streams)):
        worst = max(worst, float(np.abs(probs[s] - ref[:, 1]).max()))
        assert np.array_equal(bits(probs[s]), bits(ref[:, 1])), (s, worst)
        assert vadc_b200.segments_text(probs[s]) == Oracle().segments_text(ref[:, 1])
    assert worst == 0.0


def test_throughput_does_not_fall_when_streams_are_added():
    """No regime cliff: with the kernel family fixed per engine, the device-resident rate (chunks per second) of a call must not drop when
    the caller adds streams -- across the points where the engine changes kernel mapping (740/741 streams: LSTM wavefront launch -> one launch per layer;
    384 chunks per window: encoder CTA-per-chunk -> thread-per-token). 10 % tolerance for timer noise on these short runs."""
    N = 32
    base = [vadc_b200.synth_pcm(100 + i, CHUNK * N) for i in range(8)]
    rates = []
    sizes = (8, 11, 12, 16, 64, 74, 75, 128, 129, 296, 297, 512, 740, 741, 1023, 1024, 4096)
    e = vadc_b200.Engine(max_streams=max(sizes))
    for S in sizes:
        pcm = np.stack([base[s % 8] for s in range(S)])
        d_pcm, d_probs = e.device_alloc(pcm.nbytes), e.device_alloc(S * N * 4)
        e.h2d(d_pcm, pcm)
        best = 0.0
        for _ in range(3):
            e.reset()
            e.timer_start()
            e.run_streams_device(d_pcm, pcm.shape[1], S, N, d_probs)
            ms = e.timer_stop()
            best = max(best, S * N / ms)
        e.device_free(d_pcm)
        e.device_free(d_probs)
        rates.append(best)
    e.close()
    for (s0, r0), (s1, r1) in zip(zip(sizes, rates), list(zip(sizes, rates))[1:]):
        assert r1 >= 0.9 * r0, "rate drops from %d streams (%.0f chunks/ms) to %d streams (%.0f chunks/ms): %r" % (s0, r0, s1, r1, list(zip(sizes, rates)))
