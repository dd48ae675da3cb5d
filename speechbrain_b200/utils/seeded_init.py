"""Deterministic, order-independent synthetic weights keyed by state_dict name.

There is no network in this project (no checkpoints), so parity tests, goldens and
the benchmark all run on seeded random weights.  The reference module and ours
share state_dict keys, so the same dict loads into both.
Values depend only on (seed, key, shape) -- not on parameter creation order.
"""
import hashlib
import math

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def seeded_tensor(seed: int, key: str, shape, kind: str = "auto") -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(seed, key)
    if kind == "auto":
        if key.endswith("glob_std"):
            kind = "std"
        elif key.endswith("glob_mean"):
            kind = "mean"
        elif key.endswith(".weight") and (len(shape) == 1 or key.endswith("norm.weight")):
            kind = "gain"
        elif len(shape) >= 2:
            kind = "matrix"
        else:
            kind = "bias"
    if kind == "gain":  # LayerNorm gains around 1
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "bias":
        return 0.05 * torch.randn(shape, generator=g)
    if kind == "std":
        return 0.5 + torch.rand(shape, generator=g)
    if kind == "mean":
        return torch.randn(shape, generator=g)
    # xavier-normal-like for matrices / conv kernels
    rf = 1
    for s in shape[2:]:
        rf *= s
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return std * torch.randn(shape, generator=g)


def seeded_state_dict(module_or_state_dict, seed: int = 0):
    """Return {key: tensor} of seeded values for every float entry of a module's
    state_dict (non-float buffers and sinusoid tables are left untouched)."""
    sd = module_or_state_dict.state_dict() if hasattr(module_or_state_dict, "state_dict") else module_or_state_dict
    out = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v) or k.endswith(".pe") or k.endswith("inv_freq"):
            out[k] = v.clone()
            continue
        out[k] = seeded_tensor(seed, k, v.shape).to(v.dtype)
    return out


def seeded_asr_state(cfg, seed: int = 0):
    """Full flat state for the recipe's modules (CNN., Transformer., seq_lin., ctc_lin.) plus fixed global-CMVN
    statistics (normalize.glob_mean / glob_std), exactly as oracle/make_goldens.py gave the reference."""
    from .shapes import asr_model_shapes

    sd = {k: seeded_tensor(seed, k, shp) for k, shp in asr_model_shapes(cfg).items()}
    sd["normalize.glob_mean"] = seeded_tensor(seed, "normalize.glob_mean", (cfg["n_mels"],)) * 3.0 - 20.0
    sd["normalize.glob_std"] = seeded_tensor(seed, "normalize.glob_std", (cfg["n_mels"],)) * 8.0
    return sd


CONFORMER_LARGE = dict(name="conformer_large", sample_rate=16000, n_fft=512, win=512, hop=160, n_mels=80,
                       cnn_channels=(64, 32), input_size=640, d_model=512, nhead=8, num_encoder_layers=12,
                       num_decoder_layers=6, d_ffn=2048, vocab=5000, kernel_size=31, attention_type="RoPEMHA",
                       decoder_activation="gelu", max_length=2500)
CONFORMER_SMALL = dict(name="conformer_small", sample_rate=16000, n_fft=400, win=400, hop=160, n_mels=80,
                       cnn_channels=(64, 32), input_size=640, d_model=144, nhead=4, num_encoder_layers=12,
                       num_decoder_layers=4, d_ffn=1024, vocab=5000, kernel_size=31, attention_type="RelPosMHAXL",
                       decoder_activation="gelu", max_length=2500)
