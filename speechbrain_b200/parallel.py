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


def gather_hypotheses(tokens, n_total, world, pad=-1, max_len=None, out=None):
    """tokens [n_local, L_local] int32 -> [n_total, L] on every rank with ONE collective (``all_gather_into_tensor`` on a
    preallocated [world, per, L] buffer).

    Ranks may hold different numbers of rows (the last shards of ``shard_bounds`` can be short) and different widths
    (greedy early exit and beam search stop at rank-dependent steps), while a collective needs identical shapes: rows are
    padded to ceil(n_total / world) and columns to ``max_len`` -- pass the decode limit (max_decode_steps) to skip the
    extra all-reduce(MAX) that otherwise agrees on the width.  ``pad`` fills both."""
    if world == 1:
        return tokens
    per = (n_total + world - 1) // world
    L = tokens.shape[1]
    if max_len is None:
        width = torch.tensor([L], device=tokens.device, dtype=torch.int64)
        dist.all_reduce(width, op=dist.ReduceOp.MAX)
        max_len = int(width.item())
    if L > max_len or tokens.shape[0] > per:
        raise ValueError(f"gather_hypotheses: local block {tuple(tokens.shape)} exceeds the agreed [{per}, {max_len}]")
    if tokens.shape[0] == per and L == max_len and tokens.is_contiguous():
        buf = tokens
    else:
        buf = torch.full((per, max_len), pad, dtype=tokens.dtype, device=tokens.device)
        buf[: tokens.shape[0], :L] = tokens
    if out is None or out.shape != (world * per, max_len) or out.dtype != tokens.dtype or out.device != tokens.device:
        out = torch.empty(world * per, max_len, dtype=tokens.dtype, device=tokens.device)
    dist.all_gather_into_tensor(out, buf)
    return out[:n_total]
