"""Stream sharding across the GPUs of one box.

Streams are independent (they share only the read-only weights), so rank r of `world` owns a
contiguous block of streams, keeps their LSTM and segmenter state on its own GPU, and there is NO
collective on the data path. The only exchange is the final gather of per-stream segments to rank 0,
a few bytes per stream, done with torch.distributed object collectives (NCCL is only used by
bench.py for its barrier / max-over-ranks; this module works with any backend, `gloo` in the tests).
"""


def stream_range(n_streams, world, rank):
    """Contiguous block partition: the first n_streams % world ranks own one extra stream.
    Returns (first_stream, count) in the global stream numbering."""
    if world < 1 or not 0 <= rank < world or n_streams < 0:
        raise ValueError("bad partition request: n_streams=%r world=%r rank=%r" % (n_streams, world, rank))
    base, extra = divmod(n_streams, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(stream, n_streams, world):
    """Rank that owns global stream `stream` under stream_range."""
    if not 0 <= stream < n_streams:
        raise ValueError("stream %r outside [0, %r)" % (stream, n_streams))
    base, extra = divmod(n_streams, world)
    split = extra * (base + 1)
    return stream // (base + 1) if stream < split else extra + (stream - split) // max(base, 1)


def gather_segments(local_segments, first_stream, n_streams, group=None, dst=0):
    """local_segments: list (one entry per locally owned stream, in order) of [(start_chunk, end_chunk), ...].
    Returns, on rank `dst`, the list over ALL n_streams global streams; None elsewhere.
    Without an initialised process group (single GPU) it is the identity."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if first_stream != 0 or len(local_segments) != n_streams:
            raise ValueError("single-process gather needs all %d streams, got %d from %d" % (n_streams, len(local_segments), first_stream))
        return [list(s) for s in local_segments]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (first_stream, [list(map(tuple, s)) for s in local_segments])
    parts = [None] * world if rank == dst else None
    dist.gather_object(payload, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * n_streams
    for first, segs in parts:
        for i, s in enumerate(segs):
            if out[first + i] is not None:
                raise RuntimeError("stream %d reported by two ranks" % (first + i))
            out[first + i] = s
    missing = [i for i, s in enumerate(out) if s is None]
    if missing:
        raise RuntimeError("streams never reported: %r ..." % missing[:8])
    return out
