"""speechbrain.nnet.activations.Softmax mirror (nnet/activations.py:20-87): the recipes end their CTC branch with
``Softmax(apply_log=True)``.  Inside EncoderASR the log-softmax is fused into the CTC head kernel; called on its own this module
runs torch's (log_)softmax on the tensor's device (plumbing, not part of the timed path)."""
import torch


class Softmax(torch.nn.Module):
    def __init__(self, apply_log=False, dim=-1, reshape=True, dtype=torch.float32):
        super().__init__()
        self.apply_log, self.dim, self.dtype = apply_log, dim, dtype

    def forward(self, x):
        return torch.log_softmax(x, self.dim, dtype=self.dtype) if self.apply_log else torch.softmax(x, self.dim, dtype=self.dtype)
