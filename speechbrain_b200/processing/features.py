"""InputNormalization -- drop-in for speechbrain.processing.features.InputNormalization
(processing/features.py:1265-1551) for INFERENCE: eval-mode forward on the GPU kernels, same constructor,
same ``_save`` / ``_load`` checkpoint format ({count, glob_mean, glob_std} via torch.save)."""
import ctypes

import torch

from .._lib import check, lib, ptr, require_cuda, stream_ptr


class InputNormalization(torch.nn.Module):
    NORM_TYPES = ("global", "batch", "sentence")

    def __init__(self, mean_norm=True, std_norm=True, norm_type="global", avg_factor=None, length_dim=1,
                 update_until_epoch=2, avoid_padding_norm=False, epsilon=1e-10, device="cpu"):
        super().__init__()
        if not mean_norm:
            raise ValueError("Passing `False` for `mean_norm` is deprecated.")
        if avg_factor is not None:
            raise ValueError("Passing avg_factor is DEPRECATED (see the reference InputNormalization).")
        if norm_type == "speaker":
            raise ValueError("per-speaker normalization is deprecated.")
        elif norm_type not in self.NORM_TYPES:
            raise ValueError(f"norm_type must be one of {self.NORM_TYPES}.")
        if length_dim != 1:
            raise NotImplementedError("speechbrain_b200.InputNormalization: length_dim must be 1")
        self.std_norm, self.norm_type = std_norm, norm_type
        self.avoid_padding_norm, self.epsilon = avoid_padding_norm, epsilon
        self.device, self.length_dim = device, length_dim
        self.update_until_epoch = update_until_epoch or float("inf")
        self.glob_mean = torch.empty(0)
        self.glob_std = torch.empty(0)
        self.count = 0

    @torch.no_grad()
    def forward(self, x, lengths=None, epoch=None):
        """x [B, T, F] -> normalised [B, T, F] (eval semantics, :1404-1455)."""
        if self.training and self.norm_type == "global" and (epoch is None or epoch < self.update_until_epoch):
            raise NotImplementedError("speechbrain_b200.InputNormalization is inference-only: call .eval() "
                                      "(running-statistics updates are training-time, out of scope)")
        if self.norm_type == "batch":
            raise NotImplementedError("norm_type='batch' couples utterances; not supported on the sharded hot path")
        require_cuda(x, "InputNormalization")
        if x.dim() != 3:
            raise NotImplementedError("speechbrain_b200.InputNormalization: expected [batch, time, features]")
        x = x.float().contiguous()
        B, T, F = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            sp = stream_ptr(x.device)
            if self.norm_type == "global":
                if self.glob_mean.numel() != F:
                    raise RuntimeError("InputNormalization(global): statistics not loaded (glob_mean is empty)")
                mean = self.glob_mean.to(x.device, torch.float32).contiguous()
                std = (self.glob_std.to(x.device, torch.float32) if self.std_norm else torch.ones_like(mean)).contiguous()
                if self.avoid_padding_norm:
                    raise NotImplementedError("avoid_padding_norm with global statistics is not built")
                check(lib().sbk_input_norm_global(ptr(x), ptr(out), B, T, F, ptr(mean), ptr(std),
                                                  ctypes.c_float(self.epsilon), sp), "sbk_input_norm_global")
            else:
                rl = lengths.to(x.device, torch.float32).contiguous() if lengths is not None else None
                check(lib().sbk_input_norm_sentence(ptr(x), ptr(out), ptr(rl), B, T, F, int(bool(self.std_norm)),
                                                    int(bool(self.avoid_padding_norm)), ctypes.c_float(self.epsilon), sp),
                      "sbk_input_norm_sentence")
        return out

    # ---- checkpoint format identical to the reference (:1488-1551)
    def _statistics_dict(self):
        return {"count": self.count, "glob_mean": self.glob_mean, "glob_std": self.glob_std}

    def _load_statistics_dict(self, state):
        self.count = state["count"]
        self.glob_mean = state["glob_mean"]
        self.glob_std = state["glob_std"]
        return state

    def to(self, device):
        self.device = device
        self = super().to(device)
        self.glob_mean = self.glob_mean.to(device)
        self.glob_std = self.glob_std.to(device)
        return self

    def _save(self, path):
        torch.save(self._statistics_dict(), path)

    def _load(self, path, end_of_epoch=False):
        del end_of_epoch
        self._load_statistics_dict(torch.load(path, map_location=self.device))
