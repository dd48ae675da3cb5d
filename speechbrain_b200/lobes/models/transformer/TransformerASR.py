"""TransformerASR -- drop-in for speechbrain.lobes.models.transformer.TransformerASR.TransformerASR
(TransformerASR.py:167-675) restricted to what the Conformer ASR recipes instantiate:
encoder_module="conformer", attention_type in {"RoPEMHA", "RelPosMHAXL"}, normalize_before=True, causal=False.

Same constructor kwargs, same state_dict keys (incl. the positional buffers), ``encode()`` on the sm_100a
kernels.  ``decode()``/``forward()`` (teacher-forced training-style calls) are not part of the inference hot
path: the searchers in speechbrain_b200.decoders run the KV-cached decoder directly.
"""
import math

import torch

from ...._lib import require_cuda
from ....utils.param_tree import _Node, build_param_tree, default_init
from ....utils.shapes import transformer_asr_shapes


def _sine_table(max_len, d):
    """Transformer.py:252-303 PositionalEncoding buffer ``pe`` (1, max_len, d)."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len).unsqueeze(1).float()
    den = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.unsqueeze(0)


class TransformerASR(torch.nn.Module):
    def __init__(self, tgt_vocab, input_size, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                 d_ffn=2048, dropout=0.1, activation=torch.nn.ReLU, positional_encoding="fixed_abs_sine",
                 normalize_before=False, kernel_size=31, bias=True, encoder_module="transformer",
                 conformer_activation=None, branchformer_activation=None, attention_type="regularMHA",
                 max_length=2500, causal=None, csgu_linear_units=3072, gate_activation=None,
                 use_linear_after_conv=False, output_hidden_states=False, layerdrop_prob=0.0):
        super().__init__()
        # same argument validation as Transformer.py:141-147,206-212
        assert attention_type in ["regularMHA", "RelPosMHAXL", "hypermixing", "RoPEMHA"]
        assert positional_encoding in ["fixed_abs_sine", None]
        assert num_encoder_layers + num_decoder_layers > 0, \
            "number of encoder layers and number of decoder layers cannot both be 0!"
        if encoder_module == "conformer":
            assert normalize_before, "normalize_before must be True for Conformer"
        if causal is None:
            causal = True  # the reference warns and assumes True (TransformerASR.py:274-282)
        unsupported = []
        if encoder_module != "conformer":
            unsupported.append(f"encoder_module={encoder_module!r}")
        if attention_type not in ("RoPEMHA", "RelPosMHAXL"):
            unsupported.append(f"attention_type={attention_type!r}")
        if causal:
            unsupported.append("causal=True (streaming / chunked masks)")
        if not bias:
            unsupported.append("bias=False")
        if output_hidden_states:
            unsupported.append("output_hidden_states=True")
        if conformer_activation is not None and getattr(conformer_activation, "__name__", "") not in ("Swish", "SiLU"):
            unsupported.append("conformer_activation other than Swish")
        if d_model % nhead or d_model // nhead not in (64, 36, 32):
            unsupported.append(f"head_dim={d_model // max(nhead, 1)} (64, 36 or 32)")
        if unsupported:
            raise NotImplementedError("speechbrain_b200.TransformerASR: not built: " + ", ".join(unsupported))
        act_name = getattr(activation, "__name__", str(activation))
        if act_name not in ("GELU", "ReLU"):
            raise NotImplementedError(f"speechbrain_b200.TransformerASR: decoder activation {act_name} not built")
        self.decoder_activation = "gelu" if act_name == "GELU" else "relu"
        self.tgt_vocab, self.input_size, self.d_model, self.nhead = tgt_vocab, input_size, d_model, nhead
        self.num_encoder_layers, self.num_decoder_layers, self.d_ffn = num_encoder_layers, num_decoder_layers, d_ffn
        self.kernel_size, self.attention_type, self.max_length, self.causal = kernel_size, attention_type, max_length, causal
        self.positional_encoding_type = positional_encoding
        build_param_tree(self, transformer_asr_shapes(tgt_vocab, input_size, d_model, nhead, num_encoder_layers,
                                                      num_decoder_layers, d_ffn, kernel_size, attention_type), default_init)
        # buffers the reference keeps in its state_dict (Transformer.py:150-163)
        if attention_type == "RelPosMHAXL":
            self.positional_encoding = _Node()
            inv = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
            self.positional_encoding.register_buffer("inv_freq", inv)
        elif positional_encoding == "fixed_abs_sine":
            self.positional_encoding = _Node()
            self.positional_encoding.register_buffer("pe", _sine_table(max_length, d_model))
        self.positional_encoding_decoder = _Node()
        self.positional_encoding_decoder.register_buffer("pe", _sine_table(max_length, d_model))
        self._engine = None

    def engine_cfg(self):
        return dict(n_fft=400, hop=160, win=400, n_mels=80, cnn_channels=(64, 32), input_size=self.input_size,
                    d_model=self.d_model, nhead=self.nhead, num_encoder_layers=self.num_encoder_layers,
                    num_decoder_layers=self.num_decoder_layers, d_ffn=self.d_ffn, vocab=self.tgt_vocab,
                    kernel_size=self.kernel_size, attention_type=self.attention_type,
                    decoder_activation=self.decoder_activation, max_length=self.max_length)

    def prefixed_state(self, prefix="Transformer."):
        return {prefix + k: v for k, v in self.state_dict().items()}

    def _get_engine(self, device):
        if self._engine is None or self._engine.device != torch.device(device):
            from ....engine import AsrEngine
            self._engine = AsrEngine(self.engine_cfg(), self.prefixed_state(), device=device, parts=("encoder",))
        return self._engine

    @torch.no_grad()
    def encode(self, src, wav_len=None, pad_idx=0, dynchunktrain_config=None):
        """src [B, T, F] or [B, T, F', C] -> encoder_out [B, T, d_model] (TransformerASR.py:475-544)."""
        if dynchunktrain_config is not None:
            raise NotImplementedError("speechbrain_b200.TransformerASR: dynamic chunk training/streaming is not built")
        require_cuda(src, "TransformerASR.encode")
        if src.dim() == 4:
            bz, t, ch1, ch2 = src.shape
            src = src.reshape(bz, t, ch1 * ch2)
        if wav_len is not None and float(wav_len.max()) < 1.0 - 1e-6:
            # the reference builds its mask with width max(abs_len) and then fails to broadcast (dataio.py:836)
            raise ValueError("wav_len: the longest utterance must have relative length 1.0")
        return self._get_engine(src.device).encode_from_cnn(src, wav_len)

    def decode(self, tgt, encoder_out, enc_len=None):
        raise NotImplementedError("speechbrain_b200.TransformerASR.decode: use S2STransformer{Greedy,Beam}Searcher "
                                  "(the KV-cached decoder step replaces the reference's whole-prefix decode)")

    def forward(self, src, tgt, wav_len=None, pad_idx=0):
        raise NotImplementedError("speechbrain_b200.TransformerASR.forward (teacher-forced training pass) is out of scope; "
                                  "use encode() + a searcher")


class EncoderWrapper(torch.nn.Module):
    """TransformerASR.py:678-714: calls ``transformer.encode`` so the model can sit at the end of a Sequential."""

    def __init__(self, transformer, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.transformer = transformer

    def forward(self, x, wav_lens=None, pad_idx=0, **kwargs):
        return self.transformer.encode(x, wav_lens, pad_idx, **kwargs)
