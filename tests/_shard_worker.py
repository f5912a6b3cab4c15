"""Worker of tests/test_shard_gloo.py: one process per rank, `gloo` backend, CPU only.
Each rank owns its block of streams (vadc_b200.shard.stream_range), turns its streams' probabilities
into segments with the host state machine (the device one needs a GPU; the sharding logic is the
same), and rank 0 gathers every stream's segments. No collective touches the per-chunk data."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def probabilities(stream, nchunks):
    rng = np.random.default_rng(1000 + stream)
    walk = np.cumsum(rng.normal(0, 0.1, nchunks)) * 0.3
    return np.clip(0.45 + walk, 0, 1).astype(np.float32)


def main():
    import torch.distributed as dist

    import vadc_b200
    from vadc_b200 import shard

    n_streams, nchunks, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    first, count = shard.stream_range(n_streams, world, rank)
    local = []
    for s in range(first, first + count):
        seg = vadc_b200.StreamSegmenter()
        local.append(seg.feed(probabilities(s, nchunks)) + seg.finish())
    everything = shard.gather_segments(local, first, n_streams)
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"world": world, "segments": everything}, f)
    else:
        assert everything is None
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
