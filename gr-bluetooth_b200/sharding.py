"""Time-sharding of a capture across GPUs: in stateless mode every slot
window is a pure function of its H samples, so rank g takes a contiguous range of work() calls
plus H-1 samples of guard in front of it.  No data-path collective: each rank returns its ordered
hit list and rank 0 concatenates them in rank order (already (slot, channel, ...) sorted)."""
import numpy as np


def shard_calls(n_calls, world_size, rank):
    """Contiguous, balanced split of work() calls 0..n_calls-1 -> (first_call, count)."""
    base, rem = divmod(n_calls, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_span(first_call, count, S, H):
    """Absolute sample range [lo, hi) that calls first_call..first_call+count-1 read
    (call k reads [k*S-(H-1), k*S]); lo may be negative = zero history before the stream."""
    if count <= 0:
        return 0, 0
    return first_call * S - (H - 1), (first_call + count - 1) * S + 1


def extract_span(samples, lo, hi):
    """samples[lo:hi] with zeros where the range leaves the stream (scheduler's zero history)."""
    n = len(samples)
    out = np.zeros(hi - lo, np.complex64)
    a, b = max(lo, 0), min(hi, n)
    if b > a:
        out[a - lo:b - lo] = samples[a:b]
    return out


def gather_hits(hits, dist=None, dst=0):
    """Concatenate per-rank hit arrays on rank dst in rank order (torch.distributed, any backend)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return hits
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(hits, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    bucket = [b for b in bucket if b is not None]
    return np.concatenate(bucket) if bucket else None


def run_sharded(samples, process_fn, S, H, world_size, rank, dist=None, batch=64):
    """process_fn(iq_span, first_call, n_calls) -> structured hit array with a 'slot' field.
    Every rank processes its shard in batches; returns the merged list on rank 0."""
    n_calls = (len(samples) + S - 1) // S
    first, count = shard_calls(n_calls, world_size, rank)
    parts = []
    k = first
    while k < first + count:
        n = min(batch, first + count - k)
        lo, hi = shard_span(k, n, S, H)
        parts.append(process_fn(extract_span(samples, lo, hi), k, n))
        k += n
    mine = np.concatenate(parts) if parts else None     # an empty shard contributes nothing
    merged = gather_hits(mine, dist)
    return merged
