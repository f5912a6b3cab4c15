"""GPU: the fp32 FFT kernel behind the hybrid STFT rule of the opt-in fast family (stft.c:15-229 + misc.c:40-82) against the oracle.
STFT_HYBRID is the 8-lanes-per-frame kernel (stft_fft8_kernel.cuh). It must flag the small bins, be bit-identical to the reference on
the flagged bins, and keep the probabilities inside the 1e-4 bar on short streams."""
import numpy as np
import pytest

import vadc_b200
from oracle_lib import Oracle
from test_gpu_parity import HYB_REL, PTOL, _edge_signals, f32

pytestmark = pytest.mark.gpu
MODES = [vadc_b200.STFT_HYBRID]


@pytest.mark.parametrize("mode", MODES)
def test_fft_kernels_edge_signals_and_speech(mode):
    o = Oracle()
    e = vadc_b200.Engine(max_streams=8, stft_mode=mode)
    x = np.concatenate([_edge_signals(), f32(vadc_b200.synth_pcm(77, 1536 * 13))])   # 21 chunks: partial waves of CTAs
    st = o.run_stages(x)
    e.stft_stats(reset=True)
    mag = e.stage_stft_magnitude(x)
    total, exact = e.stft_stats(reset=True)
    assert total == x.shape[0] * 3225 and exact > 0
    assert (np.abs(mag - st["stft"]) <= HYB_REL * st["stft"]).all()
    # all-zero chunk -> exactly zero; flagged bins are bit-identical: the small ones of the speech-like chunks
    assert not mag[0].any()
    xp = np.pad(x.astype(np.float64), ((0, 0), (128, 128)), mode="reflect")
    frames = np.lib.stride_tricks.sliding_window_view(xp, 256, axis=1)[:, ::64]          # [n, 25, 256]
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(256) / 256)
    nrm = np.sqrt(((frames * hann) ** 2).sum(axis=2))                                     # ||windowed frame||_2
    small = st["stft"] < 0.5 * 0.004 * nrm[:, None, :]                                    # well inside the k_rel = 4e-3 rule
    assert small.any() and np.array_equal(mag[small], st["stft"][small])
    norm, logmag = e.stage_stft_norm(x)
    assert np.abs(norm - st["norm"]).max() < 2 * HYB_REL
    e.reset()
    assert np.abs(e.run_chunks(x) - st["out"]).max() <= PTOL
    e.close()


@pytest.mark.parametrize("mode", MODES)
def test_fft_kernels_exact_path_alone_is_bit_exact(mode):
    """k_rel = huge sends every bin through the warp-cooperative exact tree of either kernel."""
    o = Oracle()
    e = vadc_b200.Engine(stft_mode=mode, stft_k_rel=1e30)
    x = np.concatenate([_edge_signals(), f32(vadc_b200.synth_pcm(8, 1536 * 8))])
    assert np.array_equal(e.stage_stft_magnitude(x), o.run_stages(x)["stft"])
    e.close()


@pytest.mark.parametrize("mode", MODES)
def test_fft_kernels_streams_s16_vs_oracle(mode):
    """s16 entry point, several windows, more chunks than resident CTAs of the STFT grid."""
    o = Oracle()
    S, N = 70, 41
    pcm = np.stack([vadc_b200.synth_pcm(500 + s, N * 1536, kind=(0 if s % 4 else 2)) for s in range(S)])
    e = vadc_b200.Engine(max_streams=S, window_chunks=17, stft_mode=mode)
    p, out2 = e.run_streams(pcm, want_out2=True)
    for s in (0, 1, 35, 69):
        o.reset()
        ref = o.run_pcm(pcm[s])
        assert np.abs(out2[s] - ref).max() <= PTOL, s
        assert vadc_b200.segments_text(p[s]) == o.segments_text(ref[:, 1]), s
    e.close()
