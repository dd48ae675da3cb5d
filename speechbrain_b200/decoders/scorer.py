"""Scorer interfaces of the beam search -- mirrors of speechbrain.decoders.scorer for what is built on the B200 path:
``TransformerLMScorer`` (scorer.py:455-560) inside a ``ScorerBuilder`` (scorer.py:1075-1341) as a *full* scorer
(shallow fusion).  CTC / coverage / length / KenLM / RNNLM scorers and partial scorers raise."""


class TransformerLMScorer:
    def __init__(self, language_model, temperature=1.0):
        self.lm = language_model
        self.temperature = temperature


class ScorerBuilder:
    def __init__(self, weights=dict(), full_scorers=list(), partial_scorers=list(), scorer_beam_scale=2):
        if partial_scorers:
            raise NotImplementedError("speechbrain_b200.ScorerBuilder: partial scorers are not built")
        if len(full_scorers) != 1 or not isinstance(full_scorers[0], TransformerLMScorer):
            raise NotImplementedError("speechbrain_b200.ScorerBuilder: exactly one full scorer, TransformerLMScorer, is built "
                                      "(CTC / coverage / length / KenLM / RNNLM scorers are not)")
        unknown = set(weights) - {"transformerlm"}
        if unknown:
            raise ValueError(f"Weights for unavailable scorers: {sorted(unknown)}")
        if "transformerlm" not in weights:
            raise ValueError("ScorerBuilder: missing weight for 'transformerlm'")
        self.weights = {"transformerlm": float(weights["transformerlm"]), "ctc": 0.0, "length": 0.0}
        self.full_scorers = {"transformerlm": full_scorers[0]}
        self.partial_scorers = {}
        self.scorer_beam_scale = scorer_beam_scale
