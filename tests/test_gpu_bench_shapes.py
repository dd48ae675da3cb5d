"""-m gpu parity tests at the shapes bench.py runs: 10 s utterances (T = 251 = four 64-key attention blocks),
ragged lengths [1.0, 0.9, 0.6, 0.3] so that trailing key blocks are partially / fully masked, 48 greedy steps (KV-cache
positions 0..47), beam = 10 with the recipe's scorers, Conformer-small 8 x 5 s -- all against outputs of the RUNNING
REFERENCE committed under tests/golden/bench_*.pt (generator: oracle/make_goldens.py bench_*).

Bars: encoder rel-L2 <= 1e-3 (north_star); greedy tokens identical up to the first decision whose reference top-1/top-2
margin is below 5e-3 (fp16 operands move logits by ~1e-3), chosen log-probs within 2e-2; beam search: identical best
hypothesis, or -- when fp16 rounding made the search pick another near-tied hypothesis -- a hypothesis the CPU oracle
(pinned against the reference by the generator) scores within 3e-2 of what we report and no worse than the reference's best
minus 3e-2."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _cfg(g):
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, CONFORMER_SMALL
    base = CONFORMER_LARGE if g["cfg"]["name"] == "conformer_large" else CONFORMER_SMALL
    return dict(base, attention_type=g["cfg"]["attention_type"])


def _inputs(g):
    """The generator's waveform, regenerated from its seed (a checksum pins the RNG stream)."""
    B, L = g["wav_shape"]
    gen = torch.Generator().manual_seed(g["wav_seed"])
    wav = torch.randn(B, L, generator=gen)
    lens = g["wav_lens"]
    for b in range(B):
        wav[b, int(round(float(lens[b]) * L)):] = 0
    chk = float(wav.double().abs().sum())
    assert abs(chk - g["wav_checksum"]) / g["wav_checksum"] < 1e-9, "regenerated waveform differs from the golden run"
    return wav, lens


def _engine(cfg, dev, parts=("fbank", "cnn", "encoder", "decoder")):
    from speechbrain_b200.engine import AsrEngine
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    sd = seeded_asr_state(cfg, 0)
    return AsrEngine(cfg, sd, device=dev, parts=parts), sd


def _check_greedy(tag, name, pred, score, g):
    ref_tok, margin, ref_lp = g["greedy_tokens"], g["greedy_margin"], g["greedy_chosen_lp"]
    B, S = ref_tok.shape
    compared, worst, stops = 0, 0.0, []
    for b in range(B):
        for s in range(S):
            if int(pred[b, s]) != int(ref_tok[b, s]):
                assert float(margin[b, s]) < 5e-3, f"[{tag}/{name}] token mismatch at b={b} s={s}, reference margin {float(margin[b, s]):.4f}"
                stops.append((b, s))
                break
            d = abs(float(score[b, s]) - float(ref_lp[b, s]))
            worst = max(worst, d)
            assert d < 2e-2, f"[{tag}/{name}] chosen log-prob err {d} at b={b} s={s}"
            compared += 1
    print(f"[{tag}] greedy[{name}]: {compared}/{B * S} decisions compared identical, max chosen-log-prob err {worst:.2e}, "
          f"near-tie stops {stops}")
    assert compared >= 0.6 * B * S, "too few decisions comparable"


@pytest.mark.parametrize("tag", ["bench_conformer_large_rope_10s", "bench_conformer_large_relpos_10s"])
def test_bench_shape_encoder_and_greedy(dev, tag):
    """wav -> Fbank -> CMVN -> CNN -> 12 Conformer layers (multi-block flash attention, ragged key masks) -> 48 greedy steps,
    through the fused device pipeline (C ABI sbk_asr_transcribe_greedy_dev), vs the reference."""
    g = torch.load(os.path.join(GOLDEN, tag + ".pt"))
    cfg = _cfg(g)
    eng, sd = _engine(cfg, dev)
    chk = float(sum(v.double().abs().sum() for k, v in sorted(sd.items())))
    assert abs(chk - g["weight_checksum"]) / g["weight_checksum"] < 1e-9, "seeded weights differ from the golden run"
    wav, lens = _inputs(g)
    S = g["greedy_tokens"].shape[1]
    for tc_rows, name in ((1 << 30, "weight-streaming"), (1, "tcgen05")):
        eng.set_decoder_tc_min_rows(tc_rows)
        pred, score, enc, done = eng.transcribe_greedy_dev(wav.to(dev), lens.to(dev), S, 1, 2, want_enc=True)
        torch.cuda.synchronize()
        assert done == S
        enc = enc.cpu()
        assert torch.isfinite(enc).all()
        r_all = _rel(enc, g["enc_out"])
        per_utt = [_rel(enc[b, : int(g["abs_len"][b])], g["enc_out"][b, : int(g["abs_len"][b])]) for b in range(enc.shape[0])]
        print(f"[{tag}] encoder rel-L2 err {r_all:.3e} (valid frames per utterance: {['%.2e' % x for x in per_utt]}) "
              f"max abs {(enc - g['enc_out']).abs().max():.3e}")
        assert r_all < 1e-3 and max(per_utt) < 1e-3
        _check_greedy(tag, name, pred.cpu(), score.cpu(), g)


def test_bench_shape_conformer_small(dev):
    """BASELINE config 2: Conformer-small (12L / 144d / 4 heads of 36 / RelPosMHAXL / n_fft 400), 8 x 5 s ragged, wav ->
    encoder states through the fused pipeline vs the reference."""
    g = torch.load(os.path.join(GOLDEN, "bench_conformer_small_relpos_5s.pt"))
    cfg = _cfg(g)
    eng, sd = _engine(cfg, dev, parts=("fbank", "cnn", "encoder"))
    wav, lens = _inputs(g)
    enc = eng.encode_wav(wav.to(dev), lens.to(dev)).cpu()
    r = _rel(enc, g["enc_out"])
    print(f"[conformer_small 8x5s] encoder rel-L2 err {r:.3e} max abs {(enc - g['enc_out']).abs().max():.3e}")
    assert enc.shape == g["enc_out"].shape and r < 1e-3
    eng.set_poll_interval(0)  # the encode-only CUDA-graph path
    enc2 = eng.encode_wav(wav.to(dev), lens.to(dev)).cpu()
    assert torch.equal(enc, enc2)


def test_bench_shape_decode_teacher_forced(dev):
    """TransformerASR.decode(tgt, encoder_out, enc_len) (TransformerASR.py:426-473) on 4 x T=251 memories with ragged
    lengths, 48 target positions, vs the reference's decoder outputs."""
    import bench
    g = torch.load(os.path.join(GOLDEN, "bench_conformer_large_rope_10s.pt"))
    gd = torch.load(os.path.join(GOLDEN, "bench_decode_conformer_large_rope_10s.pt"))["decode"]
    cfg = _cfg(g)
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    asr = bench.build_product_asr(cfg, seeded_asr_state(cfg, 0), dev)
    tr = asr.transformer
    pred, attn = tr.decode(gd["tgt"].to(dev), g["enc_out"].to(dev), gd["enc_len"].to(dev))
    r = _rel(pred.cpu(), gd["pred"])
    print(f"decode(tgt, enc, enc_len): rel-L2 err {r:.3e} max abs {(pred.cpu() - gd['pred']).abs().max():.3e}")
    assert pred.shape == gd["pred"].shape and attn is None and r < 2e-3


BEAM_CASES = [("bench_conformer_large_rope_10s", "beam_b10_lm_ctc"), ("bench_conformer_large_rope_10s", "beam_b10_ctc_valid"),
              ("bench_decode_conformer_large_rope_10s", "beam_b10_plain_eos12"),
              ("bench_decode_conformer_large_rope_10s", "beam_b10_plain_eos16")]


@pytest.mark.parametrize("file,case", BEAM_CASES)
def test_bench_shape_beam10(dev, file, case):
    """BASELINE config 4 family: beam = 10 on T = 251 memories: [TransformerLM 0.6, CTC 0.4] (test search), [CTC] (valid
    search) for 24 steps, and scorer-less searches whose hypotheses finish gradually (up to 48 steps)."""
    g = torch.load(os.path.join(GOLDEN, "bench_conformer_large_rope_10s.pt"))
    gb = torch.load(os.path.join(GOLDEN, file + ".pt"))[case]
    _check_beam(dev, g, gb, case)


def test_beam66_recipe_width(dev):
    """beam_size = 66 (the recipe's test_beam_size, conformer_large.yaml:132) through the radix-select beam kernel, vs the
    reference on the 2 s golden."""
    g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt"))
    gb = torch.load(os.path.join(GOLDEN, "beam66_conformer_large_rope.pt"))
    _check_beam(dev, g, gb, "beam66")


def test_beam_search_with_coverage_scorer(dev):
    """ScorerBuilder(full_scorers=[CoverageScorer]) (scorer.py:788-955, penalty on the cumulative last-layer cross-attention)
    vs the reference on the 2 s golden; the weight is large enough that the result differs from the scorer-less search."""
    g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt"))
    gb = torch.load(os.path.join(GOLDEN, "beam_cov_conformer_large_rope.pt"))
    _check_beam(dev, g, gb, "coverage")


def _check_beam(dev, g, gb, case):
    import bench
    from oracle import asr_oracle as O
    from speechbrain_b200.utils.seeded_init import seeded_asr_state, seeded_state_dict
    cfg = _cfg(g)
    sd = seeded_asr_state(cfg, 0)
    sd["seq_lin.w.bias"] = sd["seq_lin.w.bias"].clone()
    sd["seq_lin.w.bias"][2] += gb["eos_bias"]
    cov = (gb["coverage_weight"], gb["coverage_threshold"]) if "coverage_weight" in gb else None
    asr = bench.build_product_asr(cfg, sd, dev, decoder="beam", beam=gb["kwargs"]["beam_size"], lm=gb["with_lm"], ctc=gb["with_ctc"],
                                  coverage=cov)
    bs = asr.mods["decoder"]
    bs.max_decode_ratio, bs.min_decode_ratio = gb["max_decode_ratio"], gb["kwargs"].get("min_decode_ratio", 0.0)
    bs.return_topk, bs.topk = True, gb["kwargs"]["beam_size"]
    enc, lens = g["enc_out"].to(dev), g["wav_lens"].to(dev)
    hyps, hlens, scores, lp = bs(enc, lens)
    if gb["with_lm"] or case.endswith("eos16"):
        # the same search with every projection (decoder AND TransformerLM step) on the tcgen05 GEMM instead of the
        # weight-streaming kernel (what wide beams / many utterances use): same hypotheses, scores within 2e-3
        bs._get_engine(dev).set_decoder_tc_min_rows(1)
        h2, l2, s2, _ = bs(enc, lens)
        bs._get_engine(dev).set_decoder_tc_min_rows(64)
        print(f"beam[{case}] tcgen05 projections: best scores {s2[:, 0].tolist()}")
        assert (s2[:, 0].cpu() - scores[:, 0].cpu()).abs().max() < 2e-3
    hyps, hlens, scores = hyps.cpu(), hlens.cpu(), scores.cpu()
    B, L = hyps.shape[0], hyps.shape[2]
    ref_h, ref_len, ref_s = gb["hyps"].long(), gb["lens"], gb["scores"]
    tol, diverged = 3e-2, []
    for b in range(B):
        n = int(torch.round(hlens[b, 0] * L)) + 1          # tokens the search stored for its best hypothesis
        n_ref = int(torch.round(ref_len[b, 0] * ref_h.shape[2])) + 1
        ours, ref = hyps[b, 0, :n].tolist(), ref_h[b, 0, :n_ref].tolist()
        assert abs(float(scores[b, 0]) - float(ref_s[b, 0])) < tol, f"best score {float(scores[b, 0])} vs reference {float(ref_s[b, 0])}"
        if ours != ref:
            diverged.append((b, ours))
    # the n-best list as a whole: sorted scores of all `beam` hypotheses track the reference's
    k = min(scores.shape[1], ref_s.shape[1])
    nbest_err = (scores[:, :k] - ref_s[:, :k]).abs().max().item()
    print(f"beam[{case}] best scores {scores[:, 0].tolist()} ref {ref_s[:, 0].tolist()}; identical best hypothesis for "
          f"{B - len(diverged)}/{B} utterances (reference top-1/top-2 gaps {(ref_s[:, 0] - ref_s[:, 1]).tolist()}); "
          f"max |n-best score - reference| over all {k} ranks {nbest_err:.2e}")
    assert nbest_err < tol
    if diverged:  # judge the near-tied alternative with the CPU oracle walked along OUR tokens
        lm = ctc = None
        if gb["with_lm"]:
            from speechbrain_b200.lobes.models.transformer.TransformerLM import TransformerLM
            lm_m = TransformerLM(vocab=5000, d_model=768, nhead=12, num_encoder_layers=12, num_decoder_layers=0, d_ffn=3072,
                                 dropout=0.0, activation=torch.nn.GELU, normalize_before=False)
            lm = dict(sd=seeded_state_dict(lm_m, seed=1), cfg=dict(d_model=768, nhead=12, num_encoder_layers=12, d_ffn=3072,
                                                                   activation="gelu"), weight=0.6, temperature=1.15)
        if gb["with_ctc"]:
            ctc = dict(w=sd["ctc_lin.w.weight"], b=sd["ctc_lin.w.bias"], weight=0.4, blank_index=0)
        idx = [b for b, _ in diverged]
        kw = {k_: v for k_, v in gb["kwargs"].items() if k_ != "beam_size"}
        ocfg = dict(g["cfg"])
        with torch.no_grad():
            o = O.beam_search(g["enc_out"][idx], g["wav_lens"][idx], sd, ocfg, sd["seq_lin.w.weight"], sd["seq_lin.w.bias"], 1, 2,
                              beam_size=1, prefix="Transformer.", lm=lm, ctc=ctc, forced=[t for _, t in diverged],
                              coverage=dict(weight=cov[0], threshold=cov[1]) if cov else None, **kw)
        for (b, toks), osc in zip(diverged, o.tolist()):
            print(f"   utterance {b}: our hypothesis ({len(toks)} tokens) scores {float(scores[b, 0]):.5f}, the oracle gives it "
                  f"{osc:.5f}; reference best {float(ref_s[b, 0]):.5f}")
            assert abs(osc - float(scores[b, 0])) < tol, "our score for our own hypothesis is off"
            assert osc > float(ref_s[b, 0]) - tol, "the search returned a clearly worse hypothesis than the reference"


def _sub_lens(g, idx):
    """Relative lengths of a subset: the oracle derives absolute lengths as round(T * rel) with T fixed, so they carry over."""
    return g["wav_lens"][idx]


def test_encoder_decoder_asr_interface(dev):
    """EncoderDecoderASR in the reference's module layout (encoder = LengthsCapableSequential(...), transformer, decoder):
    encode_batch / transcribe_batch on host and device tensors, greedy and beam decoders, one shared engine that follows
    load_state_dict."""
    import bench
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt"))
    gb = torch.load(os.path.join(GOLDEN, "beam_conformer_large_rope.pt"))["recipe"]
    cfg = _cfg(g)
    sd = seeded_asr_state(cfg, 0)
    asr = bench.build_product_asr(cfg, sd, dev)
    T = g["enc_out"].shape[1]
    n_steps = g["greedy_logits"].shape[1]
    asr.mods["decoder"].max_decode_ratio = (n_steps + 0.5) / T
    enc = asr.encode_batch(g["wav"], g["wav_lens"])
    assert _rel(enc.cpu(), g["enc_out"]) < 1.5e-3
    words_h, toks_h = asr.transcribe_batch(g["wav"].pin_memory(), g["wav_lens"])       # host tensors (C-ABI host entry)
    words_d, toks_d = asr(g["wav"].to(dev), g["wav_lens"].to(dev))                     # device tensors, forward()
    print("EncoderDecoderASR greedy tokens", toks_h, "reference", g["hyps"])
    assert toks_h == toks_d == g["hyps"] and words_h == [" ".join(map(str, h)) for h in toks_h]
    # the searcher and the interface share ONE engine; TransformerASR.encode reuses it too
    slot = asr.transformer.engine_slot(asr.mods["decoder"].engine_key())
    builds = slot.builds
    enc2 = asr.transformer.encode(g["cnn_out"].to(dev), g["wav_lens"].to(dev))
    hy, _, _, _ = asr.mods["decoder"](enc, g["wav_lens"].to(dev))
    assert slot.builds == builds and hy == g["hyps"] and _rel(enc2.cpu(), g["enc_out"]) < 1e-3
    # load_state_dict after first use must take effect (stale snapshot)
    lin = asr.mods["decoder"].fc
    bias = sd["seq_lin.w.bias"].clone()
    bias[7] += 100.0
    lin.load_state_dict({"w.weight": sd["seq_lin.w.weight"], "w.bias": bias})
    _, toks_new = asr.transcribe_batch(g["wav"].to(dev), g["wav_lens"].to(dev))
    assert slot.builds == builds + 1 and all(set(t) == {7} for t in toks_new), toks_new
    # beam decoder through the same interface, reference wiring (eos bias so that hypotheses finish)
    sd_b = dict(sd)
    sd_b["seq_lin.w.bias"] = sd["seq_lin.w.bias"].clone()
    sd_b["seq_lin.w.bias"][2] += gb["eos_bias"]
    asr_b = bench.build_product_asr(cfg, sd_b, dev, decoder="beam", beam=gb["kwargs"]["beam_size"])
    bs = asr_b.mods["decoder"]
    bs.max_decode_ratio, bs.min_decode_ratio, bs.temperature = gb["max_decode_ratio"], gb["kwargs"]["min_decode_ratio"], gb["kwargs"]["temperature"]
    bs.using_eos_threshold = gb["kwargs"]["using_eos_threshold"]
    words, toks = asr_b.transcribe_batch(g["wav"], g["wav_lens"])
    print("EncoderDecoderASR beam tokens", toks, "reference", gb["hyps"])
    assert toks == gb["hyps"]


def test_group_host_entry_matches_device(dev):
    """sbk_asr_transcribe_greedy_group_host_async (pinned host buffers, H2D/D2H inside, eager and whole-graph modes) gives the
    ids of the device-resident group call; also covers the greedy early-exit trimming with a forced EOS."""
    import bench
    from speechbrain_b200.decoders.seq2seq import greedy_exit_step
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
    cfg = dict(CONFORMER_LARGE, num_encoder_layers=2, num_decoder_layers=2)
    sd = seeded_asr_state(cfg, 0)
    asr = bench.build_product_asr(cfg, sd, dev)
    eng = asr.engine()
    gen = torch.Generator().manual_seed(3)
    B, L, S, G = 4, 48000, 16, 3
    wavs = [torch.randn(B, L, generator=gen).pin_memory() for _ in range(G)]
    lens = [torch.tensor([1.0, 0.8, 0.9, 0.5]).pin_memory() for _ in range(G)]
    asr.mods["decoder"].max_decode_ratio = (S + 0.5) / eng.num_frames(L)[1]
    ref = [torch.empty(B, S, dtype=torch.int32, device=dev) for _ in range(G)]
    eng.transcribe_greedy_group_dev([w.to(dev) for w in wavs], [l_.to(dev) for l_ in lens], S, 1, 2, ref)
    torch.cuda.synchronize()
    for poll in (8, 0):  # eager launches, then the whole call as one CUDA graph (memcpy nodes included)
        eng.set_poll_interval(poll)
        out = [torch.full((B, S), -7, dtype=torch.int32).pin_memory() for _ in range(G)]
        out_dev = [torch.full((B, S), -7, dtype=torch.int32, device=dev) for _ in range(G)]
        for _ in range(2):  # second call replays the cached graph
            asr.transcribe_batches_async(wavs, lens, out, out_dev)
            torch.cuda.synchronize()
        for g_ in range(G):
            assert torch.equal(out[g_], ref[g_].cpu()) and torch.equal(out_dev[g_], ref[g_]), f"poll={poll} batch {g_}"
    words, toks = asr.tokens_to_words(out[0])
    assert len(words) == B and all(len(t) == S for t in toks)
    # early exit: an EOS bias makes every row end at step 0..2; the reference loop breaks right after the last first-EOS
    p = torch.tensor([[5, 2, 2, 2], [2, 2, 2, 2], [7, 8, 2, 2]], dtype=torch.int32)
    assert greedy_exit_step(p, 2) == 3 and greedy_exit_step(p[:, :2], 2) == 2


def test_encoder_asr_ctc_greedy(dev):
    """EncoderASR (inference/ASR.py:176-389) with the CTC head + ctc_greedy_decode (decoders/ctc.py:335-378) vs the reference
    on the 2 s golden: log-posteriors within 2e-2, per-frame arg-max identical wherever the reference's top-1/top-2 margin
    exceeds 5e-3, and -- with the near-tie frames taken from the reference -- identical token lists (merge + blank filter)."""
    import functools

    from speechbrain_b200.decoders.ctc import ctc_greedy_decode, greedy_from_argmax
    from speechbrain_b200.inference.ASR import EncoderASR
    from speechbrain_b200.lobes.features import Fbank
    from speechbrain_b200.lobes.models.convolution import ConvolutionFrontEnd
    from speechbrain_b200.lobes.models.transformer.TransformerASR import EncoderWrapper, TransformerASR
    from speechbrain_b200.nnet.activations import Softmax
    from speechbrain_b200.nnet.containers import LengthsCapableSequential
    from speechbrain_b200.nnet.linear import Linear
    from speechbrain_b200.processing.features import InputNormalization
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt"))
    gc = torch.load(os.path.join(GOLDEN, "ctc_greedy_conformer_large_rope.pt"))
    cfg = _cfg(g)
    sd = seeded_asr_state(cfg, 0)
    fb = Fbank(n_fft=512, n_mels=80, win_length=32)
    norm = InputNormalization(norm_type="global")
    norm.glob_mean, norm.glob_std, norm.count = sd["normalize.glob_mean"], sd["normalize.glob_std"], 1
    norm.eval()
    cnn = ConvolutionFrontEnd(input_shape=(8, 10, 80), num_blocks=2, num_layers_per_block=1, out_channels=(64, 32),
                              kernel_sizes=(3, 3), strides=(2, 2), residuals=(False, False))
    cnn.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("CNN.")})
    tr = TransformerASR(input_size=640, tgt_vocab=5000, d_model=512, nhead=8, num_encoder_layers=12, num_decoder_layers=6,
                        d_ffn=2048, activation=torch.nn.GELU, encoder_module="conformer", attention_type="RoPEMHA",
                        normalize_before=True, causal=False)
    tr.load_state_dict({k[len("Transformer."):]: v for k, v in sd.items() if k.startswith("Transformer.")}, strict=False)
    ctc_lin = Linear(input_size=512, n_neurons=5000)
    for name in ("plain", "merge"):
        c = gc[name]
        bias = sd["ctc_lin.w.bias"].clone()
        bias[0] += c["bias_blank"]
        bias[17] += c["bias_tok"]
        ctc_lin.load_state_dict({"w.weight": sd["ctc_lin.w.weight"], "w.bias": bias})
        enc = LengthsCapableSequential(compute_features=fb, normalize=norm, cnn=cnn, transformer_encoder=EncoderWrapper(tr),
                                       ctc_lin=ctc_lin, log_softmax=Softmax(apply_log=True))
        asr = EncoderASR(modules=dict(encoder=enc), hparams=dict(tokenizer=None, decoding_function=functools.partial(ctc_greedy_decode, blank_id=0)),
                         run_opts={"device": str(dev)})
        lp = asr.encode_batch(g["wav"], g["wav_lens"]).cpu()
        assert lp.shape[:2] == g["enc_out"].shape[:2] and lp.shape[2] == 5000
        e = (lp[:, :, :64] - c["log_probs_head"]).abs().max().item()
        words, toks = asr.transcribe_batch(g["wav"], g["wav_lens"])
        am = lp.argmax(-1)
        T = lp.shape[1]
        bad = 0
        for b in range(lp.shape[0]):
            n = int(torch.round(g["wav_lens"][b] * T))
            strong = c["margin"][b, :n] >= 5e-3
            bad += int((am[b, :n][strong] != c["argmax"][b, :n].long()[strong]).sum())
        patched = torch.where(c["margin"] >= 5e-3, am, c["argmax"].long())
        hy = greedy_from_argmax(patched, g["wav_lens"], 0)
        print(f"EncoderASR[{name}] log-prob err {e:.2e}; strong-margin frames with another arg-max: {bad}; tokens {toks} ref {c['hyps']}")
        assert e < 2e-2 and bad == 0 and hy == c["hyps"]
        # module-by-module use of the mirror function on a CUDA tensor of log-probs
        assert ctc_greedy_decode(lp.to(dev), g["wav_lens"], blank_id=0) == toks


def test_fp16_range_scaled_weights(dev):
    """fp16 operand range, against the reference re-run on the 2 s golden with scaled weights:
    (a) FFN first layers x200, second layers / 200 (hidden activations -- the fp16-stored tensor -- in the hundreds, FFN output
        scale unchanged): still <= 1e-3 rel-L2 -- fp16 rounding is relative and the residual stream / LayerNorm statistics are
        fp32.  (x200 alone makes every FFN output dwarf the residual stream, i.e. a 24-deep NON-residual chain in which any
        rounding compounds: 2.0e-3 measured, 1.95e-3 with the exact-form SiLU -- a property of that construction, not of the
        number format);
    (b) attention in_proj x3 (logits x9, near one-hot softmax) and conv pw1 x4: the rounding of q and k (2^-11 relative)
        becomes an ABSOLUTE logit error 9x larger, so the attention branch is as accurate as fp16 (or TF32 / bf16) operands
        allow: bar 3e-3 here, printed beside the result;
    (c) FFN first layers x1e5: the FFN hidden exceeds the fp16 maximum (65504): the fp32 -> fp16 stores saturate
        (F2FP.SATFINITE), so the output stays finite (no inf -> NaN cascade)."""
    from oracle.make_goldens import scale_state
    from speechbrain_b200.engine import AsrEngine
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt"))
    gs = torch.load(os.path.join(GOLDEN, "conformer_large_rope_scaled.pt"))
    cfg = _cfg(g)
    sd = seeded_asr_state(cfg, 0)
    cnn = g["cnn_out"].reshape(g["cnn_out"].shape[0], g["cnn_out"].shape[1], -1).to(dev)
    results = []
    for name, bar in (("ffn", 1e-3), ("attn", 3e-3)):
        eng = AsrEngine(cfg, scale_state(sd, gs[name]["scales"]), device=dev, parts=("encoder",))
        enc = eng.encode_from_cnn(cnn, g["wav_lens"].to(dev)).cpu()
        r = _rel(enc, gs[name]["enc_out"])
        print(f"scaled weights [{name}] {gs[name]['scales']} (max |FFN pre-activation| {gs[name]['ffn_hidden_absmax']:.0f} in the "
              f"reference): encoder rel-L2 err {r:.3e} (bar {bar:g})")
        results.append((name, bool(torch.isfinite(enc).all()), r, bar))
    assert all(fin and r < bar for _, fin, r, bar in results), results
    eng2 = AsrEngine(cfg, scale_state(sd, dict(gs["ffn"]["scales"], ffn_w1=1e5, ffn_w2=1e-5)), device=dev, parts=("encoder",))
    enc2 = eng2.encode_from_cnn(cnn, g["wav_lens"].to(dev)).cpu()
    print(f"FFN x1e5 (hidden beyond the fp16 range): finite {bool(torch.isfinite(enc2).all())}, absmax {float(enc2.abs().max()):.2f}")
    assert torch.isfinite(enc2).all()


def test_conformer_small_decoder_greedy(dev):
    """conformer_small.yaml end to end (d_model 144, 4 heads of 36, 4 decoder layers): the decoder runs on the generic
    head-dim decode attention and the unfused-LayerNorm projections; greedy tokens / log-probs vs the reference golden."""
    g = torch.load(os.path.join(GOLDEN, "conformer_small_relpos.pt"))
    cfg = _cfg(g)
    eng, sd = _engine(cfg, dev)
    n_steps = g["greedy_logits"].shape[1]
    ref_lp = torch.log_softmax(g["greedy_logits"], -1)
    top2 = g["greedy_logits"].topk(2, -1).values
    margin = top2[..., 0] - top2[..., 1]
    ref_tok = g["greedy_logits"].argmax(-1)
    pred, score, lp, done = eng.greedy_from_enc(g["enc_out"].to(dev), g["wav_lens"].to(dev), n_steps, 1, 2, want_log_probs=True)
    pred, lp = pred.cpu(), lp.cpu()
    worst, compared = 0.0, 0
    for b in range(pred.shape[0]):
        for s in range(n_steps):
            if pred[b, s] != ref_tok[b, s]:
                assert margin[b, s] < 5e-3, f"token mismatch at b={b} s={s} with margin {margin[b, s]}"
                break
            worst = max(worst, (lp[b, s] - ref_lp[b, s]).abs().max().item())
            compared += 1
    print(f"[conformer_small] greedy tokens {pred.tolist()} ref {g['hyps']}; {compared} steps compared, max log-prob err {worst:.2e}")
    assert worst < 2e-2 and compared >= n_steps
    # and the whole path from the waveform
    p2, _, enc, _ = eng.transcribe_greedy_dev(g["wav"].to(dev), g["wav_lens"].to(dev), n_steps, 1, 2, want_enc=True)
    assert _rel(enc.cpu(), g["enc_out"]) < 1.5e-3


def test_dynamic_chunk_encode(dev):
    """TransformerASR.encode(src, wav_len, dynchunktrain_config=DynChunkTrainConfig(chunk, left)) -- chunked attention masks +
    Dynamic Chunk Convolution (the streaming-equivalent masked mode) -- vs the reference: RoPE and RelPos, finite and
    infinite left context, chunk sizes that do not divide T, ragged batch."""
    import bench
    from speechbrain_b200.utils.dynamic_chunk_training import DynChunkTrainConfig
    from speechbrain_b200.utils.seeded_init import seeded_asr_state
    gd = torch.load(os.path.join(GOLDEN, "dynchunk_conformer_large.pt"))
    asrs = {}
    for key, c in gd.items():
        att = c["attention_type"]
        g = torch.load(os.path.join(GOLDEN, "conformer_large_rope.pt" if att == "RoPEMHA" else "conformer_large_relpos.pt"))
        if att not in asrs:
            cfg = _cfg(g)
            asrs[att] = bench.build_product_asr(cfg, seeded_asr_state(cfg, 0), dev).transformer
        tr = asrs[att]
        src = g["cnn_out"].to(dev)
        enc = tr.encode(src, g["wav_lens"].to(dev), dynchunktrain_config=DynChunkTrainConfig(c["chunk_size"], c["left_context_size"])).cpu()
        r = _rel(enc, c["enc_out"])
        full = tr.encode(src, g["wav_lens"].to(dev)).cpu()  # the engine is back in full-context mode afterwards
        print(f"[dynchunk {key}] encoder rel-L2 err {r:.3e}; full-context afterwards rel {_rel(full, g['enc_out']):.3e}")
        assert r < 1e-3 and _rel(full, g["enc_out"]) < 1e-3


def test_encode_streaming_equals_masked(dev):
    """encode_streaming(chunk, context) chunk by chunk == encode(full, dynchunktrain_config) (the reference's own streaming
    test, tests/unittests/test_conformer.py, asserts exactly this equivalence): finite left context with history trimming
    (a 3-layer model so that the stream is longer than the retained window) and infinite left context, a short last chunk."""
    import bench
    from speechbrain_b200.utils.dynamic_chunk_training import DynChunkTrainConfig
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
    for att, cs, lc, n_layers in (("RoPEMHA", 8, 1, 3), ("RelPosMHAXL", 6, None, 2)):
        cfg = dict(CONFORMER_LARGE, attention_type=att, num_encoder_layers=n_layers, num_decoder_layers=1)
        tr = bench.build_product_asr(cfg, seeded_asr_state(cfg, 0), dev).transformer
        gen = torch.Generator().manual_seed(5)
        T = 20 * cs + 3
        src = torch.randn(2, T, 640, generator=gen).to(dev)
        dc = DynChunkTrainConfig(cs, lc)
        full = tr.encode(src, None, dynchunktrain_config=dc)
        ctx = tr.make_streaming_context(dc)
        outs = [tr.encode_streaming(src[:, t:t + cs].contiguous(), ctx) for t in range(0, T, cs)]
        stream = torch.cat(outs, dim=1)
        r = _rel(stream.cpu(), full.cpu())
        print(f"[streaming {att} chunk {cs} left {lc}] {len(outs)} chunks, retained history {ctx.history.shape[1]} of {T} frames, "
              f"rel diff vs masked {r:.2e}")
        assert stream.shape == full.shape and r < 3e-4  # same maths; the online-softmax key-block partition differs with the window
        if lc is not None:
            assert ctx.history.shape[1] < T


@pytest.mark.parametrize("B,L,lens,att", [(1, 16000, [1.0], "RoPEMHA"), (3, 12345, [1.0, 0.5, 0.21], "RelPosMHAXL"),
                                          (2, 1999, [1.0, 0.6], "RoPEMHA"), (5, 48000, [0.37, 1.0, 0.99, 0.5, 0.8], "RoPEMHA")])
def test_edge_shapes_vs_oracle(dev, B, L, lens, att):
    """Edge shapes through the fused wav -> ids pipeline vs the CPU oracle (pinned to the reference by the golden generator):
    a single utterance, sample counts that are not multiples of 4 (no TMA staging) or of the hop, very short audio (T = 4 encoder
    frames), batch sizes that are not powers of two, utterances padded to a fifth of the batch length."""
    from oracle import asr_oracle as O
    from speechbrain_b200.engine import AsrEngine
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
    cfg = dict(CONFORMER_LARGE, attention_type=att, num_encoder_layers=3, num_decoder_layers=2)
    sd = seeded_asr_state(cfg, 0)
    eng = AsrEngine(cfg, sd, device=dev)
    gen = torch.Generator().manual_seed(B * 1000 + L)
    wav = torch.randn(B, L, generator=gen)
    wl = torch.tensor(lens)
    for b in range(B):
        wav[b, int(round(lens[b] * L)):] = 0
    steps = 3
    pred, score, enc, done = eng.transcribe_greedy_dev(wav.to(dev), wl.to(dev), steps, 1, 2, want_enc=True)
    torch.cuda.synchronize()
    ocfg = dict(cfg, win_length=32)
    with torch.no_grad():
        feats = O.full_pipeline_features(wav, wl, sd, ocfg)
        ref = O.encode(feats, wl, sd, cfg, "Transformer.")
        T = ref.shape[1]
        out = O.greedy_search(ref, wl, sd, cfg, sd["seq_lin.w.weight"], sd["seq_lin.w.bias"], 1, 2, 0.0, (steps + 0.5) / T,
                              "Transformer.", return_logits=True)
    r = _rel(enc.cpu(), ref)
    logits = out[4]
    n = logits.shape[1]
    top2 = logits.topk(2, -1).values
    ok = True
    for b in range(B):
        for s in range(min(n, done)):
            if int(pred[b, s]) != int(logits[b, s].argmax()):
                ok = ok and float(top2[b, s, 0] - top2[b, s, 1]) < 5e-3
                break
    print(f"edge B={B} L={L} lens={lens} {att}: T={T} encoder rel-L2 {r:.3e}, steps {done}/{n}, tokens ok {ok}")
    assert enc.shape == ref.shape and torch.isfinite(enc).all() and r < 1e-3 and ok


def test_long_utterance_vs_oracle(dev):
    """Maximum-size direction: 2 x 40 s (T = 1001 encoder frames = 16 key blocks, 8 x the bench length), ragged, 2 Conformer
    layers, RoPE: encoder vs the CPU oracle."""
    from oracle import asr_oracle as O
    from speechbrain_b200.engine import AsrEngine
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
    cfg = dict(CONFORMER_LARGE, num_encoder_layers=2, num_decoder_layers=1)
    sd = seeded_asr_state(cfg, 0)
    eng = AsrEngine(cfg, sd, device=dev, parts=("fbank", "cnn", "encoder"))
    gen = torch.Generator().manual_seed(40)
    L = 640000
    wav = torch.randn(2, L, generator=gen)
    wl = torch.tensor([1.0, 0.73])
    wav[1, int(0.73 * L):] = 0
    enc = eng.encode_wav(wav.to(dev), wl.to(dev)).cpu()
    with torch.no_grad():
        ref = O.encode(O.full_pipeline_features(wav, wl, sd, dict(cfg, win_length=32)), wl, sd, cfg, "Transformer.")
    r = _rel(enc, ref)
    print(f"long utterance: T={ref.shape[1]} encoder rel-L2 {r:.3e}")
    assert enc.shape == ref.shape and r < 1e-3
