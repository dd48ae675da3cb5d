"""Pretrainer -- mirror of speechbrain.utils.parameter_transfer.Pretrainer (utils/parameter_transfer.py:33-350) for LOCAL
sources: ``collect_files(default_source)`` resolves every loadable's file inside a directory (no hub / URL fetching: there is
no network), ``load_collected()`` applies the reference's transfer hooks (utils/checkpoints.py:236-300): ``nn.Module`` ->
``load_state_dict(torch.load(path), strict=False)``; objects with a ``_load(path, end_of_epoch)`` checkpoint hook
(InputNormalization) -> that hook; ``sentencepiece.SentencePieceProcessor`` -> ``.load(path)``."""
import os

import torch

PARAMFILE_EXT = ".ckpt"


class Pretrainer:
    def __init__(self, collect_in=None, loadables=None, paths=None, custom_hooks=None, conditions=None):
        self.collect_in = collect_in
        self.loadables, self.paths, self.custom_hooks, self.conditions = {}, {}, {}, {}
        self.is_local = []
        if loadables is not None:
            self.add_loadables(loadables)
        if paths is not None:
            self.add_paths(paths)
        if custom_hooks is not None:
            self.custom_hooks.update(custom_hooks)
        if conditions is not None:
            self.conditions.update(conditions)

    def set_collect_in(self, path):
        self.collect_in = path

    def add_loadables(self, loadables):
        self.loadables.update(loadables)

    def add_paths(self, paths):
        self.paths.update(paths)

    def add_custom_hooks(self, custom_hooks):
        self.custom_hooks.update(custom_hooks)

    def is_loadable(self, name):
        if name not in self.conditions:
            return True
        cond = self.conditions[name]
        return bool(cond() if callable(cond) else cond)

    def collect_files(self, default_source=None, **kwargs):
        """name -> path: ``paths[name]`` when given (absolute, or relative to ``default_source``), else
        ``default_source/name.ckpt`` (parameter_transfer.py:188-297, local files only)."""
        out = {}
        for name in self.loadables:
            if not self.is_loadable(name):
                continue
            p = self.paths.get(name, name + PARAMFILE_EXT)
            p = str(p)
            if not os.path.isabs(p) and not os.path.exists(p):
                if default_source is None:
                    raise ValueError(f'Path not specified for "{name}", and no default_source given!')
                cand = os.path.join(str(default_source), os.path.basename(p))
                p = cand if os.path.exists(cand) else os.path.join(str(default_source), p)
            if not os.path.exists(p):
                raise FileNotFoundError(f"Pretrainer: file for loadable '{name}' not found: {p} (local sources only, no hub access)")
            out[name] = p
        self.paths.update(out)
        return out

    def load_collected(self):
        for name, obj in self.loadables.items():
            if not self.is_loadable(name):
                continue
            path = self.paths.get(name)
            if path is None or not os.path.exists(str(path)):
                raise ValueError(f'Loadable "{name}" has no collected file; call collect_files() first')
            if name in self.custom_hooks:
                self.custom_hooks[name](obj, path)
            elif hasattr(obj, "_load") and callable(obj._load):          # marked checkpoint/transfer hook
                obj._load(path, False)
            elif isinstance(obj, torch.nn.Module):                       # torch_parameter_transfer
                incompat = obj.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
                for k in incompat.missing_keys:
                    import warnings
                    warnings.warn(f"During parameter transfer to {type(obj).__name__} loading from {path}, the transferred "
                                  f"parameters did not have parameters for the key: {k}")
            elif type(obj).__name__ == "SentencePieceProcessor":         # _load_spm
                obj.load(str(path))
            else:
                raise RuntimeError(f"Don't know how to load {type(obj)}. Register default hook or add custom hook for this object.")
