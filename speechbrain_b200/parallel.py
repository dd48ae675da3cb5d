"""Multi-GPU plumbing for the hot path: utterance shards, one process per GPU, ONE collective.

Every stage of the path is independent per utterance (global CMVN is a fixed affine at inference), so ranks hold
full weight replicas and never talk until the end, where the padded int32 token matrices are all-gathered
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference does not shard evaluation at all -- every rank
decodes the full test set (core.py:657-726) -- so this is new host logic, not a mirror."""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous chunks of ceil(n/world) items (the last ranks may get fewer, possibly zero)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def shard_batch(wavs, wav_lens, rank, world):
    lo, hi = shard_bounds(wavs.shape[0], rank, world)
    return wavs[lo:hi], wav_lens[lo:hi], (lo, hi)


def gather_hypotheses(tokens, n_total, world, pad=-1):
    """tokens [n_local, L] int32 (pad = -1) -> [n_total, L] on every rank with one all_gather."""
    if world == 1:
        return tokens
    per = (n_total + world - 1) // world
    L = tokens.shape[1]
    buf = torch.full((per, L), pad, dtype=tokens.dtype, device=tokens.device)
    buf[: tokens.shape[0]] = tokens
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat(out, 0)[:n_total]
