"""Build an nn.Module tree whose state_dict keys equal a given {dotted key: shape} table, so that our
module mirrors expose the reference's checkpoint keys without mirroring its Python class hierarchy."""
import torch


class _Node(torch.nn.Module):
    pass


def build_param_tree(root: torch.nn.Module, shapes: dict, init=None):
    """Register a Parameter for every key of ``shapes`` under ``root`` (creating intermediate modules)."""
    for key, shape in shapes.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, _Node())
            mod = getattr(mod, p)
        t = torch.empty(*shape)
        if init is not None:
            init(key, t)
        mod.register_parameter(parts[-1], torch.nn.Parameter(t, requires_grad=False))
    return root


def default_init(key, t):
    """Reference-like defaults: xavier-normal matrices (TransformerASR._init_params, TransformerASR.py:672-675),
    unit LayerNorm gains, zero biases."""
    with torch.no_grad():
        if t.dim() > 1 and not key.endswith("norm.weight"):
            torch.nn.init.xavier_normal_(t)
        elif key.endswith(".weight"):
            t.fill_(1.0)
        else:
            t.zero_()
