"""world_size-2 gloo test of the multi-GPU host logic: utterances are sharded in contiguous
chunks, every rank decodes its shard independently, and ONE all-gather of the padded token matrix rebuilds the
batch in the original order.  (The decode itself is stubbed; the GPU path is covered by bench.py --gpus N.)"""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speechbrain_b200.parallel import gather_hypotheses, shard_batch


def _worker(rank, world, sync_file, out):
    os.environ["RANK"], os.environ["LOCAL_RANK"] = str(rank), str(rank)
    dist.init_process_group("gloo", init_method=f"file://{sync_file}", rank=rank, world_size=world)
    B, L = 7, 5  # deliberately not divisible by the world size
    wav = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
    lens = torch.linspace(0.5, 1.0, B)
    w, l, (lo, hi) = shard_batch(wav, lens, rank, world)
    assert w.shape[0] == hi - lo and torch.equal(w, wav[lo:hi])
    # stub decode: token matrix derived from the utterance content so order mistakes are visible
    tok = torch.full((hi - lo, L), -1, dtype=torch.int32)
    for i in range(hi - lo):
        n = 1 + int(w[i, 0].item()) % L
        tok[i, :n] = int(w[i, 0].item())
    full = gather_hypotheses(tok, B, world, max_len=L)
    assert full.shape == (B, L)
    # rank-dependent widths (greedy early exit / beam search stop at different steps on different ranks): the helper agrees
    # on the width itself (one all-reduce MAX) and pads
    w_local = L - rank  # rank 0: 5 columns, rank 1: 4
    ragged = gather_hypotheses(tok[:, :w_local].contiguous(), B, world)
    assert ragged.shape == (B, L)
    lo1 = 4  # rows of rank 1 (ceil(7 / 2) = 4 rows on rank 0)
    assert torch.equal(ragged[:lo1], full[:lo1]) and torch.equal(ragged[lo1:, : L - 1], full[lo1:, : L - 1])
    assert (ragged[lo1:, L - 1] == -1).all()
    for b in range(B):
        n = 1 + int(wav[b, 0].item()) % L
        assert full[b, :n].tolist() == [int(wav[b, 0].item())] * n and (full[b, n:] == -1).all()
    if rank == 0:
        torch.save(full, out)
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "full.pt")
        mp.spawn(_worker, args=(2, os.path.join(d, "sync"), out), nprocs=2, join=True)
        assert torch.load(out).shape == (7, 5)
