"""Utterance sharding for multi-GPU synthesis: utterances are fully independent, so a batch
is partitioned by utterance, every rank holds a full weight replica and there is NO collective on the data
path.  After synthesis the ranks exchange per-utterance output lengths (and, for benchmarks, timings) with
torch.distributed -- NCCL over NVLink on the GPU box, gloo in the CPU tests.

Host-side logic only (pure Python + torch.distributed); covered by tests/test_sharding.py with world_size 2.
"""
import torch

VOCODER_FLOP_PER_FRAME = 648_241_152
ACOUSTIC_FLOP_PER_FRAME = 63_800_000
ATTENTION_FLOP_PER_FRAME2 = 9_216


def estimate_cost(n_phonemes, frames_per_phoneme=5.0):
    """Relative cost of one utterance before its durations are known: F ~ 5 T frames;
    cost = (vocoder + acoustic) * F + attention * F^2."""
    f = float(n_phonemes) * frames_per_phoneme
    return (VOCODER_FLOP_PER_FRAME + ACOUSTIC_FLOP_PER_FRAME) * f + ATTENTION_FLOP_PER_FRAME2 * f * f


def partition_lpt(costs, world_size):
    """Longest-processing-time-first assignment: returns `world_size` lists of utterance indices.
    Deterministic (ties broken by index), every utterance assigned exactly once, and each rank's list is
    ordered longest first so that a rank's own length buckets are contiguous."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += float(costs[i])
    return shards


def bucket_by_length(indices, lengths, max_batch, max_padding_ratio=None):
    """Split a rank's utterances (sorted longest first) into batches of at most `max_batch`.  `max_padding_ratio`
    additionally bounds longest/shortest inside a batch; None (the default) does not: every kernel of the engine skips
    the tiles past an utterance's own length, so mixed lengths cost index arithmetic, not compute, while many small
    batches cost launches and host round trips (measured on B200, 64 utterances of 20..200 phonemes text->wave:
    309 ms with ratio 1.5 = 6 batches, 107 ms as one batch; gpurun_out/r2_c4_buckets.txt)."""
    batches, cur = [], []
    for i in sorted(indices, key=lambda k: (-int(lengths[k]), k)):
        too_ragged = (max_padding_ratio is not None and cur
                      and int(lengths[cur[0]]) > max_padding_ratio * max(int(lengths[i]), 1))
        if cur and (len(cur) >= max_batch or too_ragged):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def gather_output_lengths(local_indices, local_lengths, n_total, group=None, device=None):
    """All ranks learn every utterance's output length: each rank fills its own slots of an (n_total,) int64
    vector (others 0) and the vectors are summed with one all_reduce (the only collective of the path)."""
    import torch.distributed as dist
    out = torch.zeros(n_total, dtype=torch.int64, device=device)
    if len(local_indices):
        idx = torch.as_tensor(list(local_indices), dtype=torch.int64, device=device)
        out[idx] = torch.as_tensor(local_lengths, dtype=torch.int64, device=device).reshape(-1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def reduce_metrics(audio_seconds, elapsed_ms, group=None, device=None):
    """Whole-job audio seconds (SUM over ranks) and the job's time (MAX over ranks)."""
    import torch.distributed as dist
    a = torch.tensor([float(audio_seconds)], dtype=torch.float64, device=device)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(a.item()), float(t.item())
