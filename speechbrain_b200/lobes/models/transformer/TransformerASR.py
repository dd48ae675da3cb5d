"""TransformerASR -- drop-in for speechbrain.lobes.models.transformer.TransformerASR.TransformerASR
(TransformerASR.py:167-675) restricted to what the Conformer ASR recipes instantiate:
encoder_module="conformer", attention_type in {"RoPEMHA", "RelPosMHAXL"}, normalize_before=True, causal=False.

Same constructor kwargs, same state_dict keys (incl. the positional buffers), ``encode()`` on the sm_100a
kernels.  ``decode()`` / ``forward()`` run teacher-forced on the KV-cached decoder step (the searchers in
speechbrain_b200.decoders drive the same step one token at a time).
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
        if d_model % 16:
            unsupported.append("d_model not a multiple of 16")
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
        # engine slots (plain dict, not sub-modules): one shared device engine per set of modules wired to this model
        object.__setattr__(self, "_slots", {})

    def engine_cfg(self):
        return dict(n_fft=400, hop=160, win=400, n_mels=80, cnn_channels=(64, 32), input_size=self.input_size,
                    d_model=self.d_model, nhead=self.nhead, num_encoder_layers=self.num_encoder_layers,
                    num_decoder_layers=self.num_decoder_layers, d_ffn=self.d_ffn, vocab=self.tgt_vocab,
                    kernel_size=self.kernel_size, attention_type=self.attention_type,
                    decoder_activation=self.decoder_activation, max_length=self.max_length)

    def prefixed_state(self, prefix="Transformer."):
        return {prefix + k: v for k, v in self.state_dict().items()}

    def engine_slot(self, key=()):
        """The shared ``EngineSlot`` for the modules identified by ``key`` (ids of the output head / LM / CTC head wired to
        this model by a searcher); every mirror using the same modules gets the same repacked device engine."""
        from ....engine_cache import EngineSlot
        if key not in self._slots:
            slot = EngineSlot(self.engine_cfg)
            slot.sources["Transformer."] = self
            self._slots[key] = slot
        return self._slots[key]

    def invalidate_engines(self):
        """Drop every cached device engine (needed only after edits through ``param.data``; ``load_state_dict``, ``.to()``
        and in-place ops on the parameters are detected)."""
        for slot in self._slots.values():
            slot.invalidate()

    def _get_engine(self, device):
        """Engine for ``encode``: any slot that already holds the encoder (e.g. the one EncoderDecoderASR or a searcher
        built), else an encoder-only one."""
        for slot in self._slots.values():
            if "encoder" in slot.parts and slot.engine is not None:
                return slot.get(device, ("encoder",))
        return self.engine_slot().get(device, ("encoder",))

    @torch.no_grad()
    def encode(self, src, wav_len=None, pad_idx=0, dynchunktrain_config=None):
        """src [B, T, F] or [B, T, F', C] -> encoder_out [B, T, d_model] (TransformerASR.py:475-544).

        ``dynchunktrain_config`` (a ``DynChunkTrainConfig``): chunked attention + Dynamic Chunk Convolution, i.e. the masked
        evaluation mode whose outputs equal chunk-by-chunk streaming (TransformerASR.py:46-105, Conformer.py:190-313)."""
        require_cuda(src, "TransformerASR.encode")
        if src.dim() == 4:
            bz, t, ch1, ch2 = src.shape
            src = src.reshape(bz, t, ch1 * ch2)
        if wav_len is not None and float(wav_len.max()) < 1.0 - 1e-6:
            # the reference builds its mask with width max(abs_len) and then fails to broadcast (dataio.py:836)
            raise ValueError("wav_len: the longest utterance must have relative length 1.0")
        eng = self._get_engine(src.device)
        if dynchunktrain_config is None:
            return eng.encode_from_cnn(src, wav_len)
        if dynchunktrain_config.chunk_size <= 0:
            raise ValueError("DynChunkTrainConfig.chunk_size must be > 0")
        eng.set_dynchunk(dynchunktrain_config.chunk_size, dynchunktrain_config.left_context_size)
        try:
            return eng.encode_from_cnn(src, wav_len)
        finally:
            eng.set_dynchunk(0)

    # ------------------------------------------------------------------ streaming (TransformerASR.py:546-670)
    def make_streaming_context(self, dynchunktrain_config):
        """Streaming context for ``encode_streaming`` (TransformerASR.py:645-670)."""
        if dynchunktrain_config is None or dynchunktrain_config.chunk_size <= 0:
            raise ValueError("make_streaming_context needs a DynChunkTrainConfig with chunk_size > 0")
        return TransformerASRStreamingContext(dynchunktrain_config)

    @torch.no_grad()
    def encode_streaming(self, src, context):
        """Encoder output for one more chunk of ``src`` [B, chunk_size, F] (TransformerASR.py:546-643).

        The reference carries per-layer left-context caches; its outputs equal the masked full-sequence run
        (``encode(..., dynchunktrain_config)``, tests/unittests/test_conformer.py).  This implementation keeps the chunk
        *inputs* seen so far in the context and re-runs that masked encode over the window the new chunk can depend on
        (12 layers x (left context + convolution halo); everything, for an infinite left context), returning the rows of the
        new chunk: the same values, at the cost of recomputing the window instead of reusing per-layer caches."""
        require_cuda(src, "TransformerASR.encode_streaming")
        cfg = context.dynchunktrain_config
        if src.dim() == 4:
            src = src.reshape(src.shape[0], src.shape[1], -1)
        if context.history is not None and context.history.shape[1] % cfg.chunk_size != 0:
            raise ValueError("encode_streaming: only the last chunk of a stream may be shorter than chunk_size")
        hist = src if context.history is None else torch.cat([context.history, src], dim=1)
        out = self.encode(hist, None, dynchunktrain_config=cfg)[:, -src.shape[1]:].contiguous()
        if not cfg.is_infinite_left_context():  # trim to the receptive field of the next chunk, on a chunk boundary
            halo = (self.kernel_size - 1) // 2
            per_layer = max(cfg.left_context_size * cfg.chunk_size + cfg.chunk_size - 1, halo)
            keep = self.num_encoder_layers * per_layer + cfg.chunk_size
            keep = -(-keep // cfg.chunk_size) * cfg.chunk_size
            if hist.shape[1] > keep and hist.shape[1] % cfg.chunk_size == 0:
                hist = hist[:, -keep:]
        context.history = hist
        return out

    def _decoder_engine(self, device):
        """Engine for ``decode``: a searcher's slot when one is wired to this model (its engine already holds the decoder),
        else a decoder-only engine without the output head."""
        for slot in self._slots.values():
            if "seq_lin." in slot.sources:
                return slot.get(device, ("decoder",))
        return self.engine_slot().get(device, ("decoder",))

    @torch.no_grad()
    def decode(self, tgt, encoder_out, enc_len=None):
        """TransformerASR.py:426-473: tgt [n, s] token ids (bos first), encoder_out [n, T, d], enc_len [n] ABSOLUTE frame
        counts -> (prediction [n, s, d] = decoder.norm(decoder(...)), None).

        Runs teacher-forced on the KV-cached decoder step: position s attends to positions <= s (the reference's causal
        mask) and to the first enc_len frames of the memory.  The reference's second value (last layer's head-averaged
        cross-attention weights [n, s, T]) is not produced by the device decoder and is returned as None."""
        require_cuda(encoder_out, "TransformerASR.decode")
        if self.num_decoder_layers == 0:
            raise ValueError("TransformerASR.decode: the model has no decoder layers")
        out = self._decoder_engine(encoder_out.device).decode_teacher_forced(tgt.long(), encoder_out, enc_len)
        return out, None

    @torch.no_grad()
    def forward(self, src, tgt, wav_len=None, pad_idx=0):
        """TransformerASR.py:326-424 for inference: (encoder_out, decoder_out).  Target positions holding ``pad_idx`` lie
        behind every real token of their row, so the causal decoder gives the real positions the reference's values; the
        padded positions (ignored by every consumer) are computed as if the pads were tokens."""
        enc = self.encode(src, wav_len, pad_idx)
        enc_len = torch.round(wav_len.to(enc.device).float() * enc.shape[1]).int() if wav_len is not None else None
        dec, _ = self.decode(tgt, enc, enc_len)
        return enc, dec


class TransformerASRStreamingContext:
    """Mutable streaming state (TransformerASR.py:26-43): the DynChunkTrainConfig and the chunk inputs seen so far."""

    def __init__(self, dynchunktrain_config):
        self.dynchunktrain_config = dynchunktrain_config
        self.history = None


class EncoderWrapper(torch.nn.Module):
    """TransformerASR.py:678-714: calls ``transformer.encode`` so the model can sit at the end of a Sequential."""

    def __init__(self, transformer, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.transformer = transformer

    def forward(self, x, wav_lens=None, pad_idx=0, **kwargs):
        return self.transformer.encode(x, wav_lens, pad_idx, **kwargs)

    def forward_streaming(self, x, context):
        """TransformerASR.py:716-737: one chunk through ``encode_streaming``."""
        return self.transformer.encode_streaming(x, context)

    def make_streaming_context(self, *args, **kwargs):
        return self.transformer.make_streaming_context(*args, **kwargs)
