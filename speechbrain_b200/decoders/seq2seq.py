"""S2STransformerGreedySearcher -- drop-in for speechbrain.decoders.seq2seq.S2STransformerGreedySearcher
(decoders/seq2seq.py:181-367) on the KV-cached device decoder: same constructor, same 4-tuple result."""
import torch

from .._lib import require_cuda


class S2STransformerGreedySearcher(torch.nn.Module):
    def __init__(self, modules, temperature=0.0, bos_index=None, eos_index=None, min_decode_ratio=0.0,
                 max_decode_ratio=1.0, return_log_probs=True):
        super().__init__()
        if bos_index is None or eos_index is None:
            raise TypeError("bos_index and eos_index are required")
        if temperature != 0:
            raise NotImplementedError("speechbrain_b200 greedy searcher: temperature sampling is not built (temperature=0 only)")
        self.model, self.fc = modules[0], modules[1]
        self.temperature = temperature
        self.bos_index, self.eos_index = bos_index, eos_index
        self.min_decode_ratio, self.max_decode_ratio = min_decode_ratio, max_decode_ratio
        self.return_log_probs = return_log_probs  # extension: skip the (B,1,L,V) log-prob tensor when False

    def set_n_out(self):
        return self.fc.w.out_features

    def engine_key(self):
        return (id(self.fc), 0, 0)

    def engine_sources(self):
        return {"seq_lin.": self.fc}

    def _get_engine(self, device, parts=(), extra_sources=None):
        """The device engine shared with every other mirror wired to the same model / output head (engine_cache.py)."""
        src = self.engine_sources()
        src.update(extra_sources or {})
        return self.model.engine_slot(self.engine_key()).get(device, tuple(parts) + ("decoder",), src)

    @torch.no_grad()
    def forward(self, enc_states, wav_len):
        """enc_states [B, T, d], wav_len [B] relative -> (hyps, top_lengths (B,1), top_scores (B,1,L), top_log_probs (B,1,L,V))."""
        require_cuda(enc_states, "S2STransformerGreedySearcher")
        B, T, _ = enc_states.shape
        min_steps = int(T * self.min_decode_ratio)
        max_steps = int(T * self.max_decode_ratio)
        n = max(0, max_steps - min_steps)  # `for step in range(min_decode_steps, max_decode_steps)` (:226)
        eng = self._get_engine(enc_states.device)
        pred, score, lp, done = eng.greedy_from_enc(enc_states, wav_len, n, self.bos_index, self.eos_index,
                                                    want_log_probs=self.return_log_probs)
        # the device polls `has_ended.all()` only every few steps: cut back to the step at which the reference's loop
        # breaks (seq2seq.py:256), i.e. one column past the last row's first EOS
        done = greedy_exit_step(pred[:, :done], self.eos_index)
        return greedy_outputs(pred[:, :done], score[:, :done], lp[:, :done] if lp is not None else None, self.eos_index)


class S2STransformerBeamSearcher(torch.nn.Module):
    """Drop-in for speechbrain.decoders.seq2seq.S2STransformerBeamSearcher (decoders/seq2seq.py:752-804,1853-1934)
    without scorers: same constructor kwargs, same return values as ``forward(enc_states, wav_len)``.

    The decoder steps, log-softmax, EOS masking/threshold, length-normalised top-k, predecessor bookkeeping of the
    KV cache and the sequence scores run on the device (one kernel per step, ``beam_step_kernel``); the device
    records the per-step (token, predecessor, score, log-prob) history, and the finished-hypothesis bookkeeping of
    ``_update_hyps_and_scores_if_eos_token`` / ``_fill_alived_hyps_with_eos_token`` / ``_get_topk_prediction``
    (:1371-1476,1600-1630) is replayed from that history on the host (it is a few hundred integers)."""

    def __init__(self, modules, temperature=1.0, bos_index=None, eos_index=None, min_decode_ratio=0.0, max_decode_ratio=1.0,
                 beam_size=None, scorer=None, return_topk=False, topk=1, using_eos_threshold=True, eos_threshold=1.5,
                 length_normalization=True, using_max_attn_shift=False, max_attn_shift=60, minus_inf=-1e20):
        super().__init__()
        if bos_index is None or eos_index is None or beam_size is None:
            raise TypeError("bos_index, eos_index and beam_size are required")
        self.lm_scorer, self.lm_weight, self.ctc_scorer, self.ctc_weight, self.length_weight = None, 0.0, None, 0.0, 0.0
        self.coverage_weight, self.coverage_threshold = 0.0, 0.5
        if scorer is not None:
            from .scorer import ScorerBuilder
            if not isinstance(scorer, ScorerBuilder):
                raise NotImplementedError("speechbrain_b200 beam searcher: scorer must be a speechbrain_b200 ScorerBuilder")
            if length_normalization and scorer.weights["length"] > 0.0:
                raise ValueError("Length normalization is not compatible with length rewarding.")
            if "length" in scorer.full_scorers:
                self.length_weight = scorer.weights["length"]
            if "coverage" in scorer.full_scorers:
                self.coverage_weight = scorer.weights["coverage"]
                self.coverage_threshold = scorer.full_scorers["coverage"].threshold
            self.lm_scorer = scorer.full_scorers.get("transformerlm")
            self.lm_weight = scorer.weights["transformerlm"] if self.lm_scorer is not None else 0.0
            self.ctc_scorer = scorer.full_scorers.get("ctc")
            if self.ctc_scorer is not None and scorer.weights["ctc"] > 0.0:
                self.ctc_weight = scorer.weights["ctc"]
                if len({bos_index, eos_index, self.ctc_scorer.blank_index}) < 3 or self.ctc_scorer.eos_index != eos_index:
                    raise ValueError("Set blank, eos and bos to different indexes for joint ATT/CTC or CTC decoding")
        if using_max_attn_shift:
            raise NotImplementedError("speechbrain_b200 beam searcher: using_max_attn_shift is not built")
        if topk > beam_size:
            raise ValueError("topk must be <= beam_size")
        self.model, self.fc = modules[0], modules[1]
        self.temperature, self.bos_index, self.eos_index = temperature, bos_index, eos_index
        self.min_decode_ratio, self.max_decode_ratio = min_decode_ratio, max_decode_ratio
        self.beam_size, self.return_topk, self.topk = beam_size, return_topk, topk
        self.using_eos_threshold, self.eos_threshold = using_eos_threshold, eos_threshold
        self.length_normalization, self.minus_inf = length_normalization, minus_inf

    def set_n_out(self):
        return self.fc.w.out_features

    def engine_key(self):
        return (id(self.fc), id(self.lm_scorer.lm) if self.lm_scorer is not None else 0,
                id(self.ctc_scorer.ctc_fc) if self.ctc_weight > 0.0 else 0)

    def engine_sources(self):
        src = {"seq_lin.": self.fc}
        if self.lm_scorer is not None:
            src["lm."] = self.lm_scorer.lm
        if self.ctc_weight > 0.0:
            src["ctc_lin."] = self.ctc_scorer.ctc_fc
        return src

    def _get_engine(self, device, parts=(), extra_sources=None):
        """The device engine shared with every other mirror wired to the same model / heads / LM (engine_cache.py)."""
        src = self.engine_sources()
        src.update(extra_sources or {})
        parts = tuple(parts) + (("decoder", "lm") if self.lm_scorer is not None else ("decoder",))
        return self.model.engine_slot(self.engine_key()).get(device, parts, src)

    @torch.no_grad()
    def forward(self, enc_states, wav_len):
        require_cuda(enc_states, "S2STransformerBeamSearcher")
        B, T, _ = enc_states.shape
        min_steps, max_steps = int(T * self.min_decode_ratio), int(T * self.max_decode_ratio)
        if max_steps <= 0:
            raise ValueError("max_decode_ratio gives zero decoding steps")  # the reference fails on `scores` too
        hist = self._get_engine(enc_states.device).beam_from_enc(
            enc_states, wav_len, self.beam_size, max_steps, min_steps, self.bos_index, self.eos_index, self.temperature,
            self.using_eos_threshold, self.eos_threshold, self.length_normalization, self.minus_inf,
            lm_weight=self.lm_weight, lm_temperature=self.lm_scorer.temperature if self.lm_scorer is not None else 1.0,
            ctc_weight=self.ctc_weight, blank_index=self.ctc_scorer.blank_index if self.ctc_scorer is not None else -1,
            length_weight=self.length_weight, coverage_weight=self.coverage_weight, coverage_threshold=self.coverage_threshold)
        out = replay_beam_history(hist, B, self.beam_size, self.eos_index, self.topk)
        topk_hyps, topk_lengths, topk_scores, topk_log_probs = (t.to(enc_states.device) for t in out)
        if self.return_topk:
            return topk_hyps, topk_lengths, topk_scores, topk_log_probs
        best_hyps, best_lens = topk_hyps[:, 0, :], topk_lengths[:, 0]
        L = best_hyps.shape[1]
        hyps = [best_hyps[b, : int(torch.round(best_lens[b] * L))].tolist() for b in range(B)]  # undo_padding
        return hyps, best_lens, topk_scores[:, 0], topk_log_probs[:, 0, :]


def replay_beam_history(hist, B, beam_size, eos_index, topk=1):
    """Host replay of the hypothesis bookkeeping (decoders/seq2seq.py:1152-1202 sequences/log-probs, :1371-1416 EOS
    hypotheses, :1600-1630 final fill, :1418-1476 top-k) from the device's per-step history."""
    tok, pred, score, lp = hist
    n_bh = B * beam_size
    alived_seq = torch.empty(n_bh, 0, dtype=torch.long)
    alived_lp = torch.empty(n_bh, 0)
    finished = [[] for _ in range(B)]

    def add_eos_hyps(tokens, scores_):
        for index in torch.nonzero(tokens.eq(eos_index), as_tuple=True)[0].tolist():
            b = index // beam_size
            if len(finished[b]) == beam_size:
                continue
            finished[b].append((alived_seq[index, :], alived_lp[index, :], scores_[index].clone()))

    last_scores = None
    for s in range(tok.shape[0]):
        if all(len(f) == beam_size for f in finished):  # `_check_full_beams` at the top of the loop (:1668)
            break
        alived_seq = torch.cat([alived_seq.index_select(0, pred[s]), tok[s].unsqueeze(1)], dim=-1)
        alived_lp = torch.cat([alived_lp.index_select(0, pred[s]), lp[s].unsqueeze(1)], dim=-1)
        last_scores = score[s]
        add_eos_hyps(tok[s], last_scores)
    if not all(len(f) == beam_size for f in finished):
        add_eos_hyps(torch.full((n_bh,), eos_index, dtype=torch.long), last_scores)
    top_hyps, top_lp, top_scores, top_len = [], [], [], []
    for i in range(B):
        hyps, lps, scs = zip(*finished[i])
        top_hyps += hyps
        top_scores += scs
        top_lp += lps
        top_len += [len(h) for h in hyps]
    top_hyps = torch.nn.utils.rnn.pad_sequence(top_hyps, batch_first=True, padding_value=0)
    top_lp = torch.nn.utils.rnn.pad_sequence(top_lp, batch_first=True, padding_value=0)
    top_len = (torch.tensor(top_len, dtype=torch.float) - 1) / top_hyps.size(1)
    top_scores = torch.stack(top_scores, dim=0).view(B, -1)
    tk_scores, idx = top_scores.topk(topk, dim=-1)
    idx = (idx + (torch.arange(B) * beam_size).unsqueeze(1)).view(B * topk)
    return (top_hyps.index_select(0, idx).view(B, topk, -1), top_len.index_select(0, idx).view(B, topk), tk_scores,
            top_lp.index_select(0, idx).view(B, topk, -1))


def greedy_exit_step(pred, eos_index):
    """Number of steps the reference's greedy loop executes for these predictions: it breaks right after the step at which
    every row has produced EOS (decoders/seq2seq.py:249-257); without that, all of them."""
    B, L = pred.shape
    if L == 0:
        return 0
    is_eos = pred == eos_index
    if not bool(is_eos.any(1).all()):
        return L
    return int(is_eos.float().argmax(1).max()) + 1


def greedy_outputs(pred, score, log_probs, eos_index):
    """decoders/seq2seq.py:259-276,280-327: lengths = first EOS position (else L) / L; hyps exclude EOS."""
    B, L = pred.shape
    is_eos = pred == eos_index
    first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full((B,), L, device=pred.device))
    top_lengths = (first.float() / max(L, 1)).unsqueeze(1)
    first_l = first.tolist()
    pl = pred.tolist()
    hyps = [pl[b][: first_l[b]] for b in range(B)]
    top_log_probs = log_probs.unsqueeze(1) if log_probs is not None else None
    return hyps, top_lengths, score.unsqueeze(1), top_log_probs
