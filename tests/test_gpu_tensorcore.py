"""The opt-in fast family's tcgen05 kernels (layer0_tc / layer_tc / lstm_tc over vadc_b200/csrc/tc_common.cuh) against the oracle and the
reference's fixtures: within 1e-4 on short streams (their documented bar). Needs a B200."""
import numpy as np
import pytest

import vadc_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,N,window", [(70, 40, 16), (33, 21, 0), (1, 9, 4)])
def test_tensor_core_lstm_vs_oracle_and_fp32_path(S, N, window):
    """lstm_tc_kernel (tcgen05, bf16x2 split) forced on: partial stream tiles, several windows (state carried
    through global memory between launches), both outputs. Bar: 1e-4 vs the oracle; also close to the FP32 kernel."""
    from oracle_lib import Oracle
    oracle = Oracle()
    pcm = np.stack([vadc_b200.synth_pcm(4000 + s, N * 1536, kind=(0 if s % 7 else 2)) for s in range(S)])
    et = vadc_b200.Engine(max_streams=S, window_chunks=window, lstm_mode=vadc_b200.LSTM_TENSOR)
    ef = vadc_b200.Engine(max_streams=S, window_chunks=window, lstm_mode=vadc_b200.LSTM_FP32)
    pt, ot = et.run_streams(pcm, want_out2=True)
    pf, of = ef.run_streams(pcm, want_out2=True)
    assert np.abs(ot - of).max() <= 1e-4
    worst = 0.0
    for s in sorted(set([0, S // 2, S - 1]) | set(range(0, S, 9))):
        oracle.reset()
        ref = oracle.run_pcm(pcm[s])
        worst = max(worst, float(np.abs(ot[s] - ref).max()))
        assert vadc_b200.segments_text(pt[s]) == oracle.segments_text(ref[:, 1]), s
    assert worst <= 1e-4, worst
    # final state agrees with the FP32 kernel's
    for s in (0, S - 1):
        ht, ct = et.get_state(s)
        hf, cf = ef.get_state(s)
        assert np.abs(ht - hf).max() <= 1e-3 and np.abs(ct - cf).max() <= 1e-3
    et.close()
    ef.close()


# ---- tensor-core encoder layers (vadc_b200/csrc/layer0_tc_kernel.cuh, layer_tc_kernel.cuh) ------------------------------------
def _fixture_engine(overrides, **kw):
    from test_gpu_fixtures import blob
    return vadc_b200.Engine(weights=blob(overrides), max_streams=1, layer_mode=vadc_b200.LAYERS_TENSOR, **kw)


def test_tc_layers_on_the_reference_fixtures():
    """The checked-in layer fixtures (test.c atol 1e-4) through the tcgen05 layer kernel."""
    from test_gpu_fixtures import fx
    v = fx("transformer_first_layer")
    e = _fixture_engine({1 + i: v[i] for i in range(24)})
    assert np.abs(e.stage_layer(0, v[24]) - v[25]).max() < 1e-4
    e.close()
    v = fx("transformer_layers_3")
    e = _fixture_engine({49 + i: v[i] for i in range(22)})
    assert np.abs(e.stage_layer(2, v[22]) - v[23]).max() < 1e-4
    e.close()
    for name, nl in (("transformer_layers_1_2", 2), ("transformer_layers_1_2_3", 3), ("transformer_layers_1_2_3_4", 4)):
        v = fx(name)
        nw = (24, 48, 70, 94)[nl - 1]
        e = _fixture_engine({1 + i: v[i] for i in range(nw)})
        outs = e.stage_encoder(v[nw])
        assert np.abs(outs[nl - 1] - v[nw + 1]).max() < 1e-4, name
        e.close()
    v = fx("adaptive_normalization_encoder")
    e = _fixture_engine({1 + i: v[i] for i in range(94)})
    assert np.abs(e.stage_encoder(e.stage_norm(v[94]))[3] - v[95]).max() < 1e-4
    e.close()


@pytest.mark.parametrize("batch", [1, 5, 16, 17, 37, 300])
def test_tc_layers_vs_oracle_stage_tensors(batch):
    """Each encoder layer alone from the oracle's exact input of that layer, partial tiles included:
    within 2e-5 * scale of the oracle (the fp16x2 split keeps 22 bits; the FP32 kernels land at the same distance)."""
    from oracle_lib import Oracle
    o = Oracle()
    pcm = vadc_b200.synth_pcm(77, 1536 * batch)
    x = (pcm.astype(np.float32) / np.float32(32768)).reshape(-1, 1536)
    st = o.run_stages(x)
    et = vadc_b200.Engine(max_streams=1, layer_mode=vadc_b200.LAYERS_TENSOR)
    ef = vadc_b200.Engine(max_streams=1, layer_mode=vadc_b200.LAYERS_FP32)
    for layer, (k_in, k_out) in enumerate((("norm", "l1"), ("l1", "l2"), ("l2", "l3"), ("l3", "l4"))):
        ref = st[k_out]
        got = et.stage_layer(layer, st[k_in])
        scale = max(1.0, float(np.abs(ref).max()))
        err_t = float(np.abs(got - ref).max())
        err_f = float(np.abs(ef.stage_layer(layer, st[k_in]) - ref).max())
        assert err_t <= 2e-5 * scale, (layer, err_t, err_f, scale)
    et.close()
    ef.close()


@pytest.mark.parametrize("S,N,window", [(37, 23, 0), (3, 130, 50), (1, 40, 7)])
def test_tc_layers_end_to_end_vs_oracle(S, N, window):
    from oracle_lib import Oracle
    oracle = Oracle()
    pcm = np.stack([vadc_b200.synth_pcm(6000 + s, N * 1536, kind=(0 if s % 5 else 2)) for s in range(S)])
    et = vadc_b200.Engine(max_streams=S, window_chunks=window, layer_mode=vadc_b200.LAYERS_TENSOR, lstm_mode=vadc_b200.LSTM_TENSOR)
    pt, ot = et.run_streams(pcm, want_out2=True)
    worst = 0.0
    for s in sorted(set([0, S // 2, S - 1]) | set(range(0, S, 9))):
        oracle.reset()
        ref = oracle.run_pcm(pcm[s])
        worst = max(worst, float(np.abs(ot[s] - ref).max()))
        assert vadc_b200.segments_text(pt[s]) == oracle.segments_text(ref[:, 1]), s
    assert worst <= 1e-4, worst
    et.close()
