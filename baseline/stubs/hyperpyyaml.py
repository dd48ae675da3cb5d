"""Import shim so the UNMODIFIED reference (baseline/_ref, pip-installed from /root/reference with --no-deps) imports offline:
speechbrain/core.py imports two names from HyperPyYAML, which is not in the offline wheelhouse.  The bench's reference arm
builds the reference modules in Python with the recipe's kwargs and never loads a YAML file, so both raise if called."""


def resolve_references(*args, **kwargs):
    raise RuntimeError("hyperpyyaml is not installed (offline stub)")


def load_hyperpyyaml(*args, **kwargs):
    raise RuntimeError("hyperpyyaml is not installed (offline stub)")
