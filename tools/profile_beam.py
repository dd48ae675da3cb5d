"""One warm-up + one profiled beam search (BASELINE config 4: 16 x 10 s, beam 10, TransformerLM + CTC scorers) restricted to a
few steps, for `ncu --profile-from-start off` launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lm = (sys.argv[3] != "nolm") if len(sys.argv) > 3 else True
cfg = dict(CONFORMER_LARGE)
dev = torch.device("cuda:0")
asr = bench.build_product_asr(cfg, seeded_asr_state(cfg, 0), dev, decoder="beam", beam=10, lm=lm, ctc=lm)
asr.mods["decoder"].max_decode_ratio = (steps + 0.5) / 251.0
g = torch.Generator().manual_seed(1234)
wav = torch.randn(B, 160000, generator=g).to(dev)
lens = torch.ones(B, device=dev)
enc = asr.encode_batch(wav, lens)
for _ in range(2):
    asr.mods["decoder"](enc, lens)
torch.cuda.synchronize()
torch.cuda.profiler.start()
asr.mods["decoder"](enc, lens)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
