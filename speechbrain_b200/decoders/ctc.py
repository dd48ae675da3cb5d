"""CTC greedy decoding -- mirror of speechbrain.decoders.ctc.{filter_ctc_output, ctc_greedy_decode} (decoders/ctc.py:298-378)
for CUDA tensors: the per-frame arg-max runs in ``rows_logsoftmax_argmax_kernel`` (csrc/ctc_scorer.cu), the merge / blank
filter of at most T integers per utterance stays on the host like in the reference."""
from itertools import groupby

import torch


def filter_ctc_output(string_pred, blank_id=-1):
    """decoders/ctc.py:298-332: merge repetitions, then drop the blank."""
    if not isinstance(string_pred, list):
        raise ValueError("filter_ctc_out can only filter python lists")
    string_out = [i[0] for i in groupby(string_pred)]
    return list(filter(lambda elem: elem != blank_id, string_out))


def frame_argmax(probabilities):
    """[B, T, V] fp32 CUDA -> [B, T] int64 arg-max per frame (first index on ties, like torch.max)."""
    import ctypes  # noqa: F401

    from .._lib import check, lib, ptr, require_cuda, stream_ptr
    require_cuda(probabilities, "ctc_greedy_decode")
    x = probabilities.float().contiguous()
    B, T, V = x.shape
    idx = torch.empty(B, T, device=x.device, dtype=torch.int32)
    with torch.cuda.device(x.device):
        check(lib().sbk_rows_argmax_f32(ptr(x), B * T, V, ptr(idx), stream_ptr(x.device)), "sbk_rows_argmax_f32")
    return idx.long()


def ctc_greedy_decode(probabilities, seq_lens, blank_id=-1):
    """decoders/ctc.py:335-378: probabilities [B, T, V] (or log-probabilities), seq_lens relative -> list of token lists."""
    if isinstance(blank_id, int) and blank_id < 0:
        blank_id = probabilities.shape[-1] + blank_id
    batch_max_len = probabilities.shape[1]
    pred = frame_argmax(probabilities).cpu()
    return greedy_from_argmax(pred, seq_lens, blank_id, batch_max_len)


def greedy_from_argmax(pred, seq_lens, blank_id, batch_max_len=None):
    batch_max_len = batch_max_len or pred.shape[1]
    out = []
    for seq, seq_len in zip(pred.tolist(), seq_lens.cpu()):
        actual_size = int(torch.round(seq_len * batch_max_len))
        out.append(filter_ctc_output(seq[:actual_size], blank_id=blank_id))
    return out
