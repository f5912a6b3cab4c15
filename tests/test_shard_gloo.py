"""CPU: the N>1 path's host logic (stream partition + final gather of per-stream segments) with
world_size 2 and 3 over `gloo` (streams shard, no data-path collective)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import vadc_b200
from vadc_b200 import shard

HERE = os.path.dirname(os.path.abspath(__file__))


def test_stream_range_is_a_partition():
    for n in (0, 1, 2, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.stream_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
            for s in set(range(0, n, max(1, n // 13))) | ({n - 1} if n else set()):
                r = shard.owner_of(s, n, world)
                assert blocks[r][0] <= s < blocks[r][0] + blocks[r][1]
    with pytest.raises(ValueError):
        shard.stream_range(8, 2, 2)


def test_gather_without_process_group_is_identity():
    segs = [[(1, 5)], [], [(0, 3), (9, 12)]]
    assert shard.gather_segments(segs, 0, 3) == segs
    with pytest.raises(ValueError):
        shard.gather_segments(segs[:2], 1, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_segments_equal_unsharded(world, tmp_path):
    sys.path.insert(0, HERE)
    from _shard_worker import probabilities
    n_streams, nchunks = 11, 700
    out = tmp_path / "gathered.json"
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_shard_worker.py"), str(n_streams), str(nchunks), str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=240) == 0
    got = json.loads(out.read_text())
    assert got["world"] == world
    want = []
    for s in range(n_streams):
        seg = vadc_b200.StreamSegmenter()
        want.append(seg.feed(probabilities(s, nchunks)) + seg.finish())
    assert [[tuple(p) for p in s] for s in got["segments"]] == want
    assert sum(len(s) for s in want) > n_streams
