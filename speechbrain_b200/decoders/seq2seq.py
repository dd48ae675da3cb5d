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
        self._engine = None

    def set_n_out(self):
        return self.fc.w.out_features

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            from ..engine import AsrEngine
            sd = self.model.prefixed_state("Transformer.")
            sd.update({"seq_lin." + k: v for k, v in self.fc.state_dict().items()})
            self._engine = AsrEngine(self.model.engine_cfg(), sd, device=device, parts=("decoder",))
        return self._engine

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
        return greedy_outputs(pred[:, :done], score[:, :done], lp[:, :done] if lp is not None else None, self.eos_index)


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
