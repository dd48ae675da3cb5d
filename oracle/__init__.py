"""CPU oracle for the SpeechBrain ASR inference hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a plain-PyTorch (CPU, fp32) restatement of the reference algorithm
(speechbrain v1.1.0, /root/reference) for the path

    wav -> Fbank -> InputNormalization -> ConvolutionFrontEnd
        -> TransformerASR.encode (Conformer) -> S2S greedy / beam search.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline -- never from ``speechbrain_b200`` (the product fails loudly
when its CUDA library is missing; it has no CPU fallback).

Parity pinning: every function here was checked in the build container against
the *running reference implementation* (imported from /root/reference with a
two-function ``hyperpyyaml`` stub) by ``oracle/make_goldens.py``; the inputs and
reference outputs of those runs are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle_golden.py`` (CPU, no reference needed).
"""
