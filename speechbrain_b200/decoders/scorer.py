"""Scorer interfaces of the beam search -- mirrors of speechbrain.decoders.scorer for what is built on the B200 path:
``TransformerLMScorer`` (scorer.py:455-560) and ``CTCScorer`` (scorer.py:81-249, CTCPrefixScore decoders/ctc.py:46-295)
as *full* scorers of a ``ScorerBuilder`` (scorer.py:1075-1341), i.e. the recipe's ``scorer_test_search`` /
``scorer_valid_search`` (conformer_large.yaml:209-228).  Coverage / length / KenLM / RNNLM scorers and partial scorers
raise.  The scoring itself runs inside the engine's beam search (csrc/engine.cu run_beam, csrc/ctc_scorer.cu)."""


class TransformerLMScorer:
    def __init__(self, language_model, temperature=1.0):
        self.lm = language_model
        self.temperature = temperature


class CTCScorer:
    def __init__(self, ctc_fc, blank_index, eos_index, ctc_window_size=0):
        if ctc_window_size != 0:
            raise NotImplementedError("speechbrain_b200.CTCScorer: ctc_window_size != 0 is not built")
        self.ctc_fc = ctc_fc
        self.blank_index = blank_index
        self.eos_index = eos_index
        self.ctc_window_size = ctc_window_size


_NAMES = {TransformerLMScorer: "transformerlm", CTCScorer: "ctc"}
_ALL = ("ctc", "rnnlm", "transformerlm", "kenlm", "coverage", "length")


class ScorerBuilder:
    def __init__(self, weights=dict(), full_scorers=list(), partial_scorers=list(), scorer_beam_scale=2):
        assert len(weights) == len(full_scorers) + len(partial_scorers), "Weights and scorers are not matched."
        if partial_scorers:
            raise NotImplementedError("speechbrain_b200.ScorerBuilder: partial scorers are not built")
        names = []
        for impl in full_scorers:
            if type(impl) not in _NAMES:
                raise NotImplementedError(f"speechbrain_b200.ScorerBuilder: {type(impl).__name__} is not built "
                                          "(TransformerLMScorer and CTCScorer are)")
            names.append(_NAMES[type(impl)])
        if len(set(names)) != len(names):
            raise ValueError("ScorerBuilder: duplicate scorers")
        unknown = set(weights) - set(_ALL)
        if unknown:
            raise ValueError(f"Weights for unavailable scorers: {sorted(unknown)}")
        if set(weights) != set(names):
            raise ValueError(f"ScorerBuilder: weights {sorted(weights)} do not match scorers {sorted(names)}")
        if names == ["ctc", "transformerlm"]:
            pass  # order only changes the float summation order of the added scores
        self.weights = {**dict.fromkeys(_ALL, 0.0), **{k: float(v) for k, v in weights.items()}}
        self.full_scorers = dict(zip(names, full_scorers))
        self.partial_scorers = {}
        self.scorer_beam_scale = scorer_beam_scale
