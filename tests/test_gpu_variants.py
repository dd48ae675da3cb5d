"""Every diagnostic / experimental kernel variant that is compiled into libsbk.so ("Diagnostic switches")
must stay parity-green: the switches are read once per process, so each variant runs the 2 s golden in its own interpreter --
fused wav -> ids pipeline for one 32-utterance batch (weight-streaming decode) and a 3-batch group (96 live rows: tcgen05 decode
projections, and enough (row, head) items for the TMA / persistent cross-attention variants) -- and must reproduce the
reference's encoder states (1e-3 rel-L2) and greedy tokens."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SCRIPT = r"""
import json, os, sys, torch
sys.path.insert(0, %r)
from speechbrain_b200.engine import AsrEngine
from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
g = torch.load(os.path.join(%r, "tests", "golden", "conformer_large_rope.pt"))
cfg = dict(CONFORMER_LARGE)
eng = AsrEngine(cfg, seeded_asr_state(cfg, 0), device="cuda:0")
wav = g["wav"].repeat(16, 1).cuda(); lens = g["wav_lens"].repeat(16).cuda()
S = g["greedy_logits"].shape[1]
pred, score, enc, done = eng.transcribe_greedy_dev(wav, lens, S, 1, 2, want_enc=True)
rel = float((enc[:2].cpu().double() - g["enc_out"].double()).norm() / g["enc_out"].double().norm())
outs = [torch.empty(32, S, dtype=torch.int32, device="cuda") for _ in range(3)]
eng.transcribe_greedy_group_dev([wav] * 3, [lens] * 3, S, 1, 2, outs)
eng.set_poll_interval(0)  # and the whole-pipeline graph
outs2 = [torch.empty(32, S, dtype=torch.int32, device="cuda") for _ in range(3)]
eng.transcribe_greedy_group_dev([wav] * 3, [lens] * 3, S, 1, 2, outs2)
torch.cuda.synchronize()
print(json.dumps({"rel": rel, "finite": bool(torch.isfinite(enc).all()), "tok": pred[:2].cpu().tolist(),
                  "rows_equal": bool(all(torch.equal(pred[2 * i:2 * i + 2], pred[:2]) for i in range(16))),
                  "group_tok": outs[0][:2].cpu().tolist(),
                  "group_equal": bool(all(torch.equal(o, outs[0]) for o in outs) and all(torch.equal(a, b) for a, b in zip(outs, outs2)))}))
""" % (ROOT, ROOT)

VARIANTS = [{}, {"SBK_GEMM_CL4": "1"}, {"SBK_GEMM_MC": "1"}, {"SBK_GEMM_BN128": "1"}, {"SBK_GEMM_V1": "1"}, {"SBK_SILU_EXACT": "1"},
            {"SBK_CNN_UNFUSED": "1"}, {"SBK_FBANK_FR16": "1"}, {"SBK_XATT_ROWMAJOR": "1"},
            {"SBK_XATT_ROWMAJOR": "1", "SBK_DEC_XATT_TMA": "1"}, {"SBK_XATT_ROWMAJOR": "1", "SBK_DEC_XATT_PERSIST": "1"},
            {"SBK_DEC_SPLITK": "1"}, {"SBK_PDL": "1"}, {"SBK_SKINNY_MT8": "1"}, {"SBK_NO_GRAPH": "1"}, {"SBK_DEC_TC_ROWS": "1"},
            {"SBK_GEMM_PAIRS": "74"}, {"SBK_GEMM_PAIRS": "40"}, {"SBK_DEC_PRIORITY": "0"}]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: "+".join(sorted(e)) or "default")
def test_kernel_variant_parity(env):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = torch.load(os.path.join(ROOT, "tests", "golden", "conformer_large_rope.pt"))
    full_env = dict(os.environ, **env)
    p = subprocess.run([sys.executable, "-c", SCRIPT], env=full_env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    print(env, r)
    assert r["finite"] and r["rel"] < 1e-3
    assert r["tok"] == g["hyps"] and r["group_tok"] == g["hyps"] and r["rows_equal"] and r["group_equal"]


@pytest.mark.parametrize("env", [{"SBK_BEAM_SERIAL": "1"}, {"SBK_NO_GRAPH": "1"}, {"SBK_BEAM_RADIX": "1"}],
                         ids=lambda e: "+".join(sorted(e)))
def test_beam_variant_parity(env):
    """The beam-search goldens (beam 10 with every scorer combination, beam 66, coverage) with the scorer branch serialised,
    without the per-step CUDA graph, and with the radix-select beam kernel forced for every width."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_bench_shapes.py"), "-q", "-m", "gpu",
                        "-k", "beam", "-p", "no:cacheprovider"], env=dict(os.environ, **env), capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    print(p.stdout[-1500:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-1000:]
