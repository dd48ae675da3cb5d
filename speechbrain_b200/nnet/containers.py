"""LengthsCapableSequential -- mirror of speechbrain.nnet.containers.LengthsCapableSequential (nnet/containers.py:165-212)
for pre-built layers: the inference YAMLs build the ``encoder`` entry of ``EncoderDecoderASR`` with it
(compute_features -> normalize -> cnn [-> transformer_encoder]).  Shape inference / layer construction from classes
(``append(Conv2d, out_channels=...)``) is training-recipe functionality and raises."""
import inspect

import torch


def lengths_arg_exists(func):
    """nnet/containers.py helper: does ``forward`` take a ``lengths`` argument."""
    return "lengths" in inspect.signature(func).parameters


class LengthsCapableSequential(torch.nn.ModuleDict):
    def __init__(self, *layers, input_shape=None, **named_layers):
        super().__init__()
        self.takes_lengths = []
        self.input_shape = input_shape
        for layer in layers:
            self.append(layer)
        for name, layer in named_layers.items():
            self.append(layer, layer_name=name)

    def append(self, layer, *args, layer_name=None, **kwargs):
        if not isinstance(layer, torch.nn.Module) or args or kwargs:
            raise NotImplementedError("speechbrain_b200.LengthsCapableSequential: pass constructed modules "
                                      "(building layers from classes with shape inference is not built)")
        if layer_name is None:
            layer_name = str(len(self))
        elif layer_name in self:
            index = 0
            while f"{layer_name}_{index}" in self:
                index += 1
            layer_name = f"{layer_name}_{index}"
        self.add_module(layer_name, layer)
        self.takes_lengths.append(lengths_arg_exists(layer.forward))

    def forward(self, x, lengths=None):
        for layer, give_lengths in zip(self.values(), self.takes_lengths):
            x = layer(x, lengths=lengths) if give_lengths else layer(x)
            if isinstance(x, tuple):
                x = x[0]
        return x
