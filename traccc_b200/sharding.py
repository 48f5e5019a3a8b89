"""Event sharding across GPUs (host logic). Events are independent: rank r
of W processes events r, r+W, r+2W, ... — no data-path collective, NCCL/gloo only carries
the end-of-run bookkeeping (counts, max elapsed time)."""
from __future__ import annotations


def events_of_rank(n_events: int, rank: int, world: int) -> list[int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return list(range(rank, n_events, world))


def stream_of_event(local_index: int, n_streams: int) -> int:
    """Round-robin of a rank's events over its algorithm instances / CUDA streams
    (one instance per stream, like one full_chain_algorithm per host thread upstream)."""
    return local_index % max(1, n_streams)
