"""ConvolutionFrontEnd -- drop-in for speechbrain.lobes.models.convolution.ConvolutionFrontEnd
(lobes/models/convolution.py:116-203) for the configuration every Conformer recipe uses:
2 blocks x 1 Conv2d(3x3, stride 2, reflect 'same' padding) + LayerNorm + LeakyReLU, no residuals.
Same constructor and state_dict keys; other configurations raise (no CPU fallback)."""
import torch

from ..._lib import require_cuda
from ...utils.param_tree import build_param_tree, default_init
from ...utils.shapes import cnn_frontend_shapes


class ConvolutionFrontEnd(torch.nn.Module):
    def __init__(self, input_shape, num_blocks=3, num_layers_per_block=5, out_channels=(128, 256, 512),
                 kernel_sizes=(3, 3, 3), strides=(1, 2, 2), dilations=(1, 1, 1), residuals=(True, True, True),
                 conv_module=None, activation=torch.nn.LeakyReLU, norm="LayerNorm", dropout=0.1, conv_bias=True,
                 padding="same", conv_init=None):
        super().__init__()
        ok = (num_blocks == 2 and num_layers_per_block == 1 and tuple(kernel_sizes[:2]) == (3, 3)
              and tuple(strides[:2]) == (2, 2) and not any(residuals[:2]) and tuple(dilations[:2]) == (1, 1)
              and conv_module is None and activation is torch.nn.LeakyReLU and norm == "LayerNorm"
              and conv_bias and padding == "same" and tuple(out_channels[:2]) == (64, 32))
        if not ok:
            raise NotImplementedError(
                "speechbrain_b200.ConvolutionFrontEnd: only the Conformer recipes' front-end is built "
                "(num_blocks=2, num_layers_per_block=1, out_channels=(64, 32), 3x3, stride 2, no residuals)")
        self.n_mels = int(input_shape[-1])
        self.out_channels = tuple(out_channels[:2])
        build_param_tree(self, cnn_frontend_shapes(self.n_mels, self.out_channels), default_init)
        object.__setattr__(self, "_slot", None)

    def _engine_cfg(self):
        f2 = ((self.n_mels - 1) // 2 + 1 - 1) // 2 + 1
        return dict(n_fft=400, hop=160, win=400, n_mels=self.n_mels, cnn_channels=self.out_channels,
                    input_size=f2 * self.out_channels[1], d_model=64, nhead=1, num_encoder_layers=0,
                    num_decoder_layers=0, d_ffn=64, vocab=8, attention_type="RoPEMHA")

    def _get_engine(self, device):
        """Stand-alone use (module-by-module pipelines): a CNN-only engine, rebuilt when the parameters change."""
        if self._slot is None:
            from ...engine_cache import EngineSlot
            object.__setattr__(self, "_slot", EngineSlot(self._engine_cfg))
        return self._slot.get(device, ("cnn",), {"CNN.": self})

    @torch.no_grad()
    def forward(self, x):
        """x [B, T, F] -> [B, T', F', C] (channels-last, like the reference)."""
        require_cuda(x, "ConvolutionFrontEnd")
        if x.dim() != 3:
            raise NotImplementedError("ConvolutionFrontEnd: expected [batch, time, features]")
        out = self._get_engine(x.device).cnn(x)
        B, T2, _ = out.shape
        return out.reshape(B, T2, -1, self.out_channels[1])
