"""One warm-up + one profiled pass of the bench workload (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speechbrain_b200.engine import AsrEngine  # noqa: E402
from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state  # noqa: E402

att = sys.argv[1] if len(sys.argv) > 1 else "RoPEMHA"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
G = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # > 1: G batches encoded one by one and decoded together (bench default 8)
cfg = dict(CONFORMER_LARGE, attention_type=att)
eng = AsrEngine(cfg, seeded_asr_state(cfg, 0), device="cuda:0")
g = torch.Generator().manual_seed(1234)
wav = torch.randn(B, 160000, generator=g).cuda()
lens = torch.ones(B).cuda()
outs = [torch.empty(B, steps, dtype=torch.int32, device="cuda") for _ in range(G)]


def run():
    if G == 1:
        eng.transcribe_greedy_dev(wav, lens, steps, 1, 2)
    else:
        eng.transcribe_greedy_group_dev([wav] * G, [lens] * G, steps, 1, 2, outs)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
