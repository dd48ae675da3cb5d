"""EncoderDecoderASR -- drop-in for speechbrain.inference.ASR.EncoderDecoderASR (inference/ASR.py:35-173)
for the Conformer encoder-decoder recipes: ``encode_batch`` / ``transcribe_batch`` with the same signatures,
running wav -> token ids as ONE fused device pipeline (C ABI sbk_asr_transcribe_greedy_*).

The reference builds its modules from a HyperPyYAML file (``from_hparams``); hyperpyyaml is not available
offline, so this class is constructed from module objects:

    asr = EncoderDecoderASR(modules=dict(compute_features=Fbank(...), normalize=InputNormalization(...),
                                         CNN=ConvolutionFrontEnd(...), Transformer=TransformerASR(...),
                                         seq_lin=Linear(...), decoder=S2STransformerGreedySearcher(...)),
                            hparams=dict(tokenizer=sentencepiece_processor_or_None))
"""
import torch

from ..decoders.seq2seq import S2STransformerGreedySearcher, greedy_outputs


class EncoderDecoderASR(torch.nn.Module):
    HPARAMS_NEEDED = ["tokenizer"]
    MODULES_NEEDED = ["compute_features", "normalize", "CNN", "Transformer", "seq_lin", "decoder"]

    def __init__(self, modules, hparams=None, run_opts=None):
        super().__init__()
        for k in self.MODULES_NEEDED:
            if k not in modules:
                raise ValueError(f"Need modules['{k}']")
        self.mods = torch.nn.ModuleDict(modules)
        self.hparams = dict(hparams or {})
        self.tokenizer = self.hparams.get("tokenizer")
        self.device = torch.device((run_opts or {}).get("device", "cuda:0"))
        self._engine = None
        if not isinstance(self.mods["decoder"], S2STransformerGreedySearcher):
            raise NotImplementedError("EncoderDecoderASR: the fused pipeline takes a S2STransformerGreedySearcher decoder")
        if self.mods["normalize"].norm_type != "global":
            raise NotImplementedError("EncoderDecoderASR: fused pipeline needs InputNormalization(norm_type='global')")

    @classmethod
    def from_hparams(cls, *args, **kwargs):
        raise NotImplementedError("speechbrain_b200.EncoderDecoderASR.from_hparams needs hyperpyyaml + network access; "
                                  "construct from module objects instead (see class docstring)")

    def engine(self):
        if self._engine is None:
            from ..engine import AsrEngine
            fb, tr = self.mods["compute_features"], self.mods["Transformer"]
            cfg = tr.engine_cfg()
            cfg.update(n_fft=fb.n_fft, hop=fb.hop_length, win=fb.win_length, n_mels=fb.n_mels, sample_rate=fb.sample_rate,
                       cnn_channels=self.mods["CNN"].out_channels)
            sd = tr.prefixed_state("Transformer.")
            sd.update({"CNN." + k: v for k, v in self.mods["CNN"].state_dict().items()})
            sd.update({"seq_lin." + k: v for k, v in self.mods["seq_lin"].state_dict().items()})
            n = self.mods["normalize"]
            sd["normalize.glob_mean"] = n.glob_mean.float().cpu()
            sd["normalize.glob_std"] = (n.glob_std if n.std_norm else torch.ones_like(n.glob_mean)).float().cpu()
            self._engine = AsrEngine(cfg, sd, device=self.device)
        return self._engine

    def _steps(self, n_samples):
        dec = self.mods["decoder"]
        _, T = self.engine().num_frames(n_samples)
        return max(0, int(T * dec.max_decode_ratio) - int(T * dec.min_decode_ratio))

    @torch.no_grad()
    def encode_batch(self, wavs, wav_lens):
        """inference/ASR.py:100-128: wavs [B, L] (+ relative lengths) -> encoder states [B, T, d]."""
        wavs = wavs.float().to(self.device)
        wav_lens = wav_lens.to(self.device)
        dec = self.mods["decoder"]
        _, _, enc, _ = self.engine().transcribe_greedy_dev(wavs, wav_lens, 0, dec.bos_index, dec.eos_index, want_enc=True)
        return enc

    @torch.no_grad()
    def transcribe_batch(self, wavs, wav_lens):
        """inference/ASR.py:131-169: -> (predicted_words list[str], predicted_tokens list[list[int]])."""
        dec = self.mods["decoder"]
        n = self._steps(wavs.shape[1])
        if wavs.is_cuda:
            pred, score, _, done = self.engine().transcribe_greedy_dev(wavs, wav_lens.to(wavs.device), n, dec.bos_index,
                                                                       dec.eos_index)
            pred = pred[:, :done].cpu()
        else:  # host buffers: H2D / D2H inside the C-ABI call
            pred, done = self.engine().transcribe_greedy_host(wavs, wav_lens, n, dec.bos_index, dec.eos_index)
            pred = pred[:, :done]
        hyps, _, _, _ = greedy_outputs(pred, torch.zeros_like(pred, dtype=torch.float32), None, dec.eos_index)
        if self.tokenizer is not None:
            words = [self.tokenizer.decode_ids(h) for h in hyps]
        else:
            words = [" ".join(map(str, h)) for h in hyps]
        return words, hyps

    def forward(self, wavs, wav_lens):
        return self.transcribe_batch(wavs, wav_lens)
