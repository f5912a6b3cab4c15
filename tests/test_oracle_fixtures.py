"""CPU: the oracle restatement against every golden fixture the reference checks in for this path
(testdata/*.testtensor copied to tests/golden/; tolerances are the reference's own: test.c atol 1e-4,
decoder_test 1e-10 -> checked at 1e-7 here because libm expf differs by <1 ulp from the torch CPU
kernel that wrote the fixture)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import ROOT, Oracle
from testtensor_io import dump_testtensor, load_list, load_testtensor

G = os.path.join(ROOT, "tests", "golden")
ATOL = 1e-4


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


@pytest.fixture(scope="module")
def o():
    return Oracle()


def fx(name):
    return load_list(os.path.join(G, name + ".testtensor"))


def test_dw_conv_129(o):
    x, w, b, exp = fx("dw_conv_129")
    out = np.zeros_like(exp)
    o.lib.so_dw_conv(_p(_f(x)), 129, x.shape[1], _p(_f(w)), _p(_f(b)), _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_pw_conv_129_16(o):
    x, w, b, exp = fx("pw_conv_129_16")
    out = np.zeros_like(exp)
    o.lib.so_pw_conv(_p(_f(x)), 129, x.shape[1], _p(_f(w)), _p(_f(b)), 16, _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_first_layer_conv_block(o):
    dw_w, dw_b, pw_w, pw_b, pr_w, pr_b, x, exp = fx("first_layer_conv_block")
    out = np.zeros_like(exp)
    o.lib.so_conv_block(_p(_f(x)), 129, x.shape[1], 1, _p(_f(dw_w)), _p(_f(dw_b)), _p(_f(pw_w)), _p(_f(pw_b)), _p(_f(pr_w)),
                        _p(_f(pr_b)), 16, _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_decoder(o):
    x, w, b, exp = fx("decoder_test")
    out = np.zeros(2, np.float32)
    o.lib.so_decoder(_p(_f(x)), 1, 64, 7, _p(_f(w)), _p(_f(b)), 2, _p(out))
    assert np.abs(out - exp.reshape(-1)).max() < 1e-7


def test_softmax(o):
    x, exp = fx("softmax_test")
    y = _f(x).copy()
    o.lib.so_softmax_rows(_p(y), 100, 100)
    assert np.abs(y - exp).max() < ATOL


def test_layernorm(o):
    x, w, b, exp = fx("layernorm_test")
    out = np.zeros_like(exp)
    o.lib.so_layer_norm(_p(_f(x)), 25, 16, _p(_f(w)), _p(_f(b)), _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_batchnorm(o):
    x, mean, var, w, b, exp = fx("batchnorm_test")
    out = np.zeros_like(exp)
    o.lib.so_batch_norm(_p(_f(x)), 50, 16, 13, _p(_f(mean)), _p(_f(var)), _p(_f(w)), _p(_f(b)), _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_dual_head_attention(o):
    x, qw, qb, pw, pb, exp = fx("dual_head_attention_test")
    out = np.zeros_like(exp)
    o.lib.so_attention(_p(_f(x)), 25, 16, _p(_f(qw)), _p(_f(qb)), _p(_f(pw)), _p(_f(pb)), _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_transformer_block(o):
    t = load_testtensor(os.path.join(G, "transformer_block_test_16_16_48.testtensor"))
    v = list(t.values())
    # fixture order: attention (4), norm1 (2), norm2 (2), linear1 (2), linear2 (2), input, result (test.c:1143)
    attn, n1, n2, l1, l2 = v[0:4], v[4:6], v[6:8], v[8:10], v[10:12]
    order = attn + n1 + l1 + l2 + n2  # fill_transformer_weights order (tensor.h:131-142)
    arr = (C.c_void_p * 12)(*[_p(_f(a)).value for a in order])
    keep = [_f(a) for a in order]
    arr = (C.c_void_p * 12)(*[k.ctypes.data for k in keep])
    x, exp = v[12], v[13]
    out = np.zeros_like(exp)
    o.lib.so_transformer_block(_p(_f(x)), 16, 25, arr, _p(out))
    assert np.abs(out - exp).max() < ATOL


def _layer_w(tensors, cin, T, cout, stride, has_proj, x, o):
    keep = [_f(a) for a in tensors]
    arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
    out = np.zeros((cout, 1 + (T - 1) // stride), np.float32)
    o.lib.so_transformer_layer_w(_p(_f(x)), cin, T, cout, stride, has_proj, arr, _p(out))
    return out


def test_transformer_first_layer(o):
    v = fx("transformer_first_layer")
    out = _layer_w(v[:24], 129, 25, 16, 2, 1, v[24][0], o)
    assert np.abs(out - v[25][0]).max() < ATOL


def test_transformer_layers_3(o):
    v = fx("transformer_layers_3")
    out = _layer_w(v[:22], 32, 7, 32, 1, 0, v[22][0], o)
    assert np.abs(out - v[23][0]).max() < ATOL


@pytest.mark.parametrize("name,nl", [("transformer_layers_1_2", 2), ("transformer_layers_1_2_3", 3), ("transformer_layers_1_2_3_4", 4),
                                     ("adaptive_normalization_encoder", 4)])
def test_cumulative_layers(o, name, nl):
    v = fx(name)
    nw = (24, 48, 70, 94)[nl - 1]
    x, exp = v[nw], v[nw + 1]
    cfg = ((129, 25, 16, 2, 1), (16, 13, 32, 2, 1), (32, 7, 32, 1, 0), (32, 7, 64, 1, 1))
    cur = _f(x[0])
    if name == "adaptive_normalization_encoder":
        cur = cur.copy()
        o.lib.so_adaptive_norm(_p(cur), 1, 129, 25)
    first = 0
    for l in range(nl):
        cin, T, cout, s, proj = cfg[l]
        n = 24 if proj else 22
        cur = _layer_w(v[first:first + n], cin, T, cout, s, proj, cur, o)
        first += n
    assert np.abs(cur - exp[0]).max() < ATOL


def test_real_weight_fixtures_match_shipped_weights():
    """The encoder weights inside the two real-weight fixtures are the shipped weights."""
    w = load_list(os.path.join(ROOT, "vadc_b200", "weights", "silero_v31_16k.testtensor"))
    for name in ("transformer_layers_1_2_3_4", "adaptive_normalization_encoder"):
        v = fx(name)
        for i in range(94):
            assert np.array_equal(v[i], w[1 + i]), (name, i)


def test_adaptive_audio_normalization(o):
    x, exp = fx("adaptive_audio_normalization_test")
    y = _f(x).copy()
    o.lib.so_adaptive_norm(_p(y), 5, 129, 25)
    assert np.abs(y - exp).max() < ATOL


def test_lstm(o):
    x, h0, c0, w, b, exp = fx("lstm_nito_reference_randn")
    out = np.zeros((11, 64), np.float32)
    o.lib.so_lstm_seq(_p(_f(x)), 7, 64, _p(_f(h0)), _p(_f(c0)), _p(_f(w)), _p(_f(b)), 2, _p(out))
    assert np.abs(out - exp).max() < ATOL


def test_testtensor_roundtrip():
    a = [("x", np.arange(6, dtype=np.float32).reshape(2, 3)), ("yy", np.ones(5, np.float32))]
    t = load_testtensor(dump_testtensor(a))
    assert list(t.keys()) == ["x", "yy"] and np.array_equal(t["x"], a[0][1])
