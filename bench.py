#!/usr/bin/env python
"""bench.py -- RTFx (audio-seconds / second) of the ASR inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Conformer-L
(conformer_large.yaml: 12L/512d/8h RoPEMHA encoder, 6L decoder, vocab 5000, n_fft=512), random init,
batch = 32 x 10 s @ 16 kHz synthetic per GPU, Fbank -> global CMVN -> CNN front-end -> encoder -> greedy
search pinned to 48 decode steps (random weights never emit EOS).  Weak scaling: every GPU
gets its own 32 utterances; the one exchange is an NCCL all-gather of the token matrix (inside the timed
region).  One "step" = one pass of the whole path over one batch.

value  : device-timed throughput, wav already resident in HBM (per-step CUDA events, L2 flushed between steps).
e2e    : the same through the host-buffer C-ABI call (pinned host wav -> H2D -> ... -> D2H token ids).
roofline: the dominant kernel (gemm_tc2_kernel, 2-CTA tcgen05) timed live with CUDA events in a separate pass.
cpu_baseline: the CPU oracle (a restatement of the reference algorithm, no KV cache) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
UTT_SECONDS = 10.0
BATCH = 32
DECODE_STEPS = 48
BOS, EOS = 1, 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


def encoder_flops_per_utt(cfg, T):
    """per frame per layer 8*d*ffn + 6d^2 + 2d^2 + 4*T*d + 4d^2 + 2*K*d + 2d^2, x T x layers,
    + CNN + input linear."""
    d, f, K = cfg["d_model"], cfg["d_ffn"], cfg["kernel_size"]
    per = 8 * d * f + 6 * d * d + 2 * d * d + 4 * T * d + 4 * d * d + 2 * K * d + 2 * d * d
    return per * T * cfg["num_encoder_layers"] + 23.1e6 + 185.0e6 + 2 * T * cfg["input_size"] * d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def usable_threads():
    """Host threads for the CPU legs: the cores this process may run on, capped at 16 -- the reference path is
    thousands of small ops per decode step and slows down badly beyond that (measured: 128 threads = 60x slower)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 16))


def synth_batch(B, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, int(SAMPLE_RATE * UTT_SECONDS), generator=g), torch.ones(B)


def cpu_oracle_rtfx(cfg, sd, B, steps, threads=None):
    """Time the CPU oracle (reference algorithm restated in torch: no KV cache, full-prefix decode) on B utterances."""
    import torch

    from oracle import asr_oracle as O
    if threads:
        torch.set_num_threads(threads)
    wav, lens = synth_batch(B, 1234)
    ocfg = dict(cfg, win_length=int(cfg["win"] * 1000 / SAMPLE_RATE))
    t0 = time.perf_counter()
    with torch.no_grad():
        feats = O.full_pipeline_features(wav, lens, sd, ocfg)
        enc = O.encode(feats, lens, sd, cfg, "Transformer.")
        T = enc.shape[1]
        O.greedy_search(enc, lens, sd, cfg, sd["seq_lin.w.weight"], sd["seq_lin.w.bias"], BOS, EOS, 0.0,
                        (steps + 0.5) / T, "Transformer.")
    dt = time.perf_counter() - t0
    return B * UTT_SECONDS / dt, dt


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port; the Python reference cannot travel to the GPU box)
    on this box's host cores, all threads, same workload definition, bounded sample per step."""
    import torch

    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CONFORMER_LARGE)
    sd = seeded_asr_state(cfg, 0)
    cores = usable_threads()
    torch.set_num_threads(cores)
    sample_B = 4
    vals = []
    t_start = time.perf_counter()
    for i in range(args.warmup_ref + args.steps_ref):
        v, dt = cpu_oracle_rtfx(cfg, sd, sample_B, DECODE_STEPS)
        if i >= args.warmup_ref:
            vals.append((v, dt))
        if vals and time.perf_counter() - t_start > 150.0:  # keep the whole arm within a few minutes on slow hosts
            break
    args.steps_ref = len(vals)
    value = sample_B * UTT_SECONDS * len(vals) / sum(dt for _, dt in vals)
    sample = f"{sample_B} x 10 s utterances per step (of the 32-utterance batch), encode + {DECODE_STEPS} greedy steps"
    line = {"metric": "audio-sec/sec (RTFx) Conformer-L ASR, batch=32x10s@16kHz", "impl": "reference", "value": value,
            "unit": "audio-sec/sec", "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
            "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "conformer_large (RoPEMHA) encode+greedy, 48 decode steps, 10 s utterances",
                       "global_batch": sample_B, "timing": "host wall clock (CPU)"},
            "cpu_baseline": {"value": value, "unit": "audio-sec/sec", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "audio-sec/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--attention", default="RoPEMHA", choices=["RoPEMHA", "RelPosMHAXL"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=4, help="batches in flight per GPU (engine clones on their own streams)")
    ap.add_argument("--group", type=int, default=16, help="batches whose decode is coalesced into one greedy loop (engine-level)")
    ap.add_argument("--decode-steps", type=int, default=48, help="diagnostic: override the pinned 48 greedy steps")
    ap.add_argument("--fuse-dec-ln", type=int, default=1, help="1: decoder LayerNorm fused into projections (latency mode)")
    args = ap.parse_args()
    args.steps_ref = max(1, args.steps)  # K steps of a 4-utterance sample each (~0.8 s on 16 cores), capped at 150 s
    args.warmup_ref = 1 if args.warmup > 0 else 0
    global DECODE_STEPS
    DECODE_STEPS = args.decode_steps
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from speechbrain_b200 import _lib
    from speechbrain_b200.engine import AsrEngine
    from speechbrain_b200.utils.seeded_init import CONFORMER_LARGE, seeded_asr_state

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps

    cfg = dict(CONFORMER_LARGE, attention_type=args.attention)
    sd = seeded_asr_state(cfg, 0)
    eng = AsrEngine(cfg, sd, device=dev)
    wav_host, lens_host = synth_batch(BATCH, 1234 + rank)
    wav_host, lens_host = wav_host.pin_memory(), lens_host.pin_memory()
    wav_dev, lens_dev = wav_host.to(dev), lens_host.to(dev)
    L = wav_host.shape[1]
    T_f, T = eng.num_frames(L)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gathered = [torch.empty(BATCH, DECODE_STEPS, dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None
    lib = _lib.lib()

    # ---- lanes: independent GROUPS of batches in flight on their own streams; weights shared, workspaces private.
    # A group = G batches of 32 x 10 s: each batch is encoded on its own, the G*32 hypotheses are decoded together.
    G = max(1, min(args.group, -(-K // max(1, args.lanes))))  # keep every lane busy when K is small
    n_calls = -(-K // G)
    sizes = [K // n_calls + (1 if i < K % n_calls else 0) for i in range(n_calls)]  # exactly K batches, balanced groups
    G = max(sizes)
    NL = max(1, min(args.lanes, n_calls))
    lanes = [eng] + [eng.clone() for _ in range(NL - 1)]
    for e in lanes:
        e.set_decoder_ln_fusion(args.fuse_dec_ln)
        e.set_poll_interval(0)  # exactly DECODE_STEPS steps, never block the host (random weights never emit EOS)
    streams = [torch.cuda.Stream(device=dev) for _ in range(NL)]
    wavs = [[wav_dev.clone() for _ in range(G)] for _ in range(NL)]
    lens = [[lens_dev.clone() for _ in range(G)] for _ in range(NL)]
    preds = [[torch.empty(BATCH, DECODE_STEPS, dtype=torch.int32, device=dev) for _ in range(G)] for _ in range(NL)]
    scores = [torch.empty(BATCH, DECODE_STEPS, dtype=torch.float32, device=dev) for _ in range(NL)]
    preds_host = [[torch.empty(BATCH, DECODE_STEPS, dtype=torch.int32).pin_memory() for _ in range(G)] for _ in range(NL)]

    def step_dev(i):
        ln, g = i % NL, sizes[i % n_calls]
        with torch.cuda.stream(streams[ln]):
            lanes[ln].transcribe_greedy_group_dev(wavs[ln][:g], lens[ln][:g], DECODE_STEPS, BOS, EOS, preds[ln][:g])
            if world > 1:
                for g_ in range(g):
                    dist.all_gather(gathered, preds[ln][g_])  # the path's only collective: final hypothesis gather

    # host->device copies go through ONE copy stream in call order (a FIFO over PCIe): the first group's 16 batches arrive
    # after 1/4 of the time it takes when the 4 lanes' copies share the link, so the pipeline fills 3 groups earlier
    h2d_stream = torch.cuda.Stream(device=dev)

    def step_host(i):
        ln, g = i % NL, sizes[i % n_calls]
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_stream(streams[ln])  # the lane's previous group no longer reads these buffers
            for g_ in range(g):  # H2D of every batch's wav / lengths from pinned host memory, inside the timed region
                wavs[ln][g_].copy_(wav_host, non_blocking=True)
                lens[ln][g_].copy_(lens_host, non_blocking=True)
        with torch.cuda.stream(streams[ln]):
            streams[ln].wait_stream(h2d_stream)
            lanes[ln].transcribe_greedy_group_dev(wavs[ln][:g], lens[ln][:g], DECODE_STEPS, BOS, EOS, preds[ln][:g])
            for g_ in range(g):  # D2H of every batch's token ids
                preds_host[ln][g_].copy_(preds[ln][g_], non_blocking=True)
                if world > 1:
                    dist.all_gather(gathered, preds[ln][g_])

    def timed(fn, n):
        """n steps between ONE pair of CUDA events; every lane stream starts after the start event and the stop
        event is recorded after all lane streams have drained."""
        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for s_ in streams + [h2d_stream]:
            s_.wait_event(e0)
        t_host = time.perf_counter()
        for i in range(n):
            fn(i)
        timed.host_ms = (time.perf_counter() - t_host) * 1e3 / max(n, 1)
        for s_ in streams:
            cur.wait_stream(s_)
        e1.record(cur)
        e1.synchronize()
        return e0.elapsed_time(e1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for rep in range(max(1, -(-W // K))):  # warm-up = the timed schedule itself (captures every (lane, group size) graph)
        for i in range(n_calls):
            step_dev(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.sbk_launch_count()
    ms_dev = timed(step_dev, n_calls)
    host_enqueue_ms = timed.host_ms * n_calls / K
    launches = lib.sbk_launch_count() - launches0  # kernels of libsbk.so inside the K-step timed region (graph nodes included)
    barrier()
    for i in range(n_calls):
        step_host(i)
    barrier()
    ms_host = timed(step_host, n_calls)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # single batch in flight (latency view): per-step events, 256 MiB L2 flush between steps
    ms_single = 0.0
    for i in range(min(K, 5)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.transcribe_greedy_dev(wav_dev, lens_dev, DECODE_STEPS, BOS, EOS, pred=preds[0][0], score=scores[0])
        e1.record()
        e1.synchronize()
        ms_single += e0.elapsed_time(e1) / min(K, 5)

    t = torch.tensor([ms_dev, ms_host], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_host = float(t[0]), float(t[1])
    audio = world * BATCH * UTT_SECONDS * K
    value = audio / (ms_dev / 1e3)
    e2e = audio / (ms_host / 1e3)

    def step_dev1():
        eng.set_poll_interval(8)  # per-kernel launches (no whole-pipeline graph) so the GEMM launches can be event-timed
        eng.transcribe_greedy_dev(wav_dev, lens_dev, DECODE_STEPS, BOS, EOS, pred=preds[0][0], score=scores[0])
        eng.set_poll_interval(0)

    # ---- roofline leg: dominant kernel = gemm_tc_kernel (tcgen05 GEMM), timed live per launch with CUDA events
    roof = None
    if rank == 0:
        hbm, tf_sus, tf_burst, src = peaks()
        torch.cuda.synchronize()
        lib.sbk_gemm_profile_enable(1)
        step_dev1()
        torch.cuda.synchronize()
        import ctypes
        n = ctypes.c_int()
        ms = ctypes.c_double()
        fl = ctypes.c_double()
        _lib.check(lib.sbk_gemm_profile_read(ctypes.byref(n), ctypes.byref(ms), ctypes.byref(fl)), "gemm_profile_read")
        lib.sbk_gemm_profile_enable(0)
        ach = fl.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0.0
        enc_fl = BATCH * encoder_flops_per_utt(cfg, T)
        roof = {"bound": "tensor",
                "kernel": "gemm_tc2_kernel<MODE,ACT,EW> (2-CTA tcgen05.mma cta_group::2 kind::f16, 256x256x64 tiles, fp16 in / "
                          "fp32 acc in TMEM) -- all encoder / cross-K,V GEMM launches of one 32 x 10 s batch",
                "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the 6 GEMMs of one encoder layer in
                # the `ncu --set full` capture profiles/r1f_gemm_ncu_full_summary.csv (reads 21.9 MB, writes 0.03 MB: the
                # outputs stay in the 126 MB L2 for the next kernel; algorithmic operand + output bytes of the same 6
                # launches average 42.4 MB)
                "traffic": 21.9e6,
                "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": n.value, "gemm_ms_per_step": ms.value, "gemm_flops_per_step": fl.value,
                "gemm_share_of_gpu_time_per_step": ms.value / (ms_dev / K),
                "encoder_flops_per_step": enc_fl,
                "encoder_roofline_rtfx": BATCH * UTT_SECONDS / (enc_fl / (tf_sus * 1e12)),
                "frac_of_encoder_roofline": (value / world) / (BATCH * UTT_SECONDS / (enc_fl / (tf_sus * 1e12)))}
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = usable_threads()
        v, dt = cpu_oracle_rtfx(cfg, sd, 16, DECODE_STEPS, threads=cores)
        cpu_base = {"value": v, "unit": "audio-sec/sec", "cores": cores, "kind": "port",
                    "sample": f"16 x 10 s utterances (half a batch), encode + {DECODE_STEPS} greedy steps, {dt:.1f} s wall"}
    if rank == 0:
        line = {"metric": "audio-sec/sec (RTFx) Conformer-L ASR, batch=32x10s@16kHz", "value": value, "unit": "audio-sec/sec",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate/residual/softmax",
                "data": "synthetic",
                "config": {"workload": f"conformer_large.yaml ({args.attention}) random init: Fbank+CMVN+CNN+12L Conformer encode "
                                       f"+ greedy {DECODE_STEPS} steps (6L decoder, KV-cached), 32 x 10 s per GPU",
                           "global_batch": world * BATCH, "utt_seconds": UTT_SECONDS, "enc_frames": T,
                           "parallelism": f"dp{world} (utterance shards, one NCCL all-gather of token ids)",
                           "lanes": NL, "decode_group": G, "decoder_ln_fused": bool(args.fuse_dec_ln), "single_lane_ms_per_step": ms_single, "host_enqueue_ms_per_step": host_enqueue_ms,
                           "l2": "no flush inside the K-step bracket: per-step working set (0.25 GB weights + 0.3 GB "
                                 "activations/KV per lane) exceeds the 126 MB L2; single_lane_ms_per_step is flushed (256 MiB) per step",
                           "timing": f"one CUDA-event pair around K steps (= {n_calls} group calls of {G} batches), {NL} groups in flight "
                                     f"on {NL} streams; max over ranks"},
                "e2e": {"value": e2e, "unit": "audio-sec/sec", "ms_per_step": ms_host / K,
                        "h2d_bytes_per_step": BATCH * L * 4 + BATCH * 4, "d2h_bytes_per_step": BATCH * DECODE_STEPS * 4},
                "single_batch": {"value": BATCH * UTT_SECONDS / (ms_single / 1e3), "unit": "audio-sec/sec", "ms_per_step": ms_single,
                                 "note": "one batch in flight, no decode coalescing, L2 flushed before every step (latency view)"},
                "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(K, 1), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
