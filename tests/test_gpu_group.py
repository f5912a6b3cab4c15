"""GPU: the multi-device stream scheduler in C (silero_b200_group_*, vadc_b200/csrc/group.c). On a one-GPU box the "devices" are
several engines on the same GPU: the sharding, the fan-out of partial stream ranges, the in-place gather into the caller's arrays and
the per-device state are exercised all the same. With more than one GPU visible the same tests use distinct devices."""
import os
import subprocess

import numpy as np
import pytest

import vadc_b200
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
CHUNK = 1536
CLI = os.path.join(os.path.dirname(vadc_b200.LIB_PATH), "vadc_b200_cli")


def devices(n):
    import ctypes
    try:
        cnt = ctypes.c_int()
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(cnt))
        have = max(1, cnt.value)
    except OSError:
        have = 1
    return [i % have for i in range(n)]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_group_equals_one_engine_and_the_oracle():
    """11 streams over 3 devices (blocks of 4, 4, 3), fed in two calls + a closing call; then a partial range that straddles two
    devices: probabilities bit-identical to a single engine and the oracle, segments identical, state carried per device."""
    S, N = 11, 40
    pcm = np.stack([vadc_b200.synth_pcm(700 + s, N * CHUNK) for s in range(S)])
    g = vadc_b200.Group(devices(3), S)
    assert g.info() == {"ndevices": 3, "streams_per_device": 4, "max_streams": S}
    g.segments_configure()
    p1, s1 = g.run_streams_segments(pcm[:, :25 * CHUNK].copy())
    p2, s2 = g.run_streams_segments(pcm[:, 25 * CHUNK:].copy(), end_of_stream=True)
    probs = np.concatenate([p1, p2], 1)
    e = vadc_b200.Engine(max_streams=S)
    pe = e.run_streams(pcm)
    e.close()
    assert np.array_equal(bits(probs), bits(pe))
    o = Oracle()
    for s in (0, 3, 4, 7, 8, 10):
        o.reset()
        ref = o.run_pcm(pcm[s])[:, 1]
        assert np.array_equal(bits(probs[s]), bits(ref))
        fmt = vadc_b200.StreamSegmenter()
        text = "".join(fmt.format(x) for x in s1[s] + s2[s])
        assert text == o.segments_text(ref)
    # partial range over the boundary between device 0 and device 1 (streams 2..5), from a fresh state
    g.reset()
    p3, _ = g.run_streams_segments(pcm[2:6, :10 * CHUNK].copy(), first_stream=2)
    assert np.array_equal(bits(p3), bits(pe[2:6, :10]))
    g.close()


def test_group_rejects_bad_ranges():
    g = vadc_b200.Group(devices(2), 6)
    with pytest.raises(vadc_b200.EngineError):
        g.run_streams_segments(np.zeros((4, 4 * CHUNK), np.int16), first_stream=4)
    g.close()


def test_cli_devices_option(tmp_path):
    """vadc_b200_cli --devices a,b: five files over two devices, output identical to the one-device run."""
    paths = []
    for i, n in enumerate((300, 120, 300, 77, 200)):
        p = tmp_path / ("f%d.s16le" % i)
        vadc_b200.synth_pcm(40 + i, n * CHUNK).tofile(p)
        paths.append(str(p))
    one = subprocess.run([CLI] + paths, capture_output=True)
    two = subprocess.run([CLI, "--devices", ",".join(str(d) for d in devices(2))] + paths, capture_output=True)
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout == two.stdout and one.stdout.count(b"#") == 5
