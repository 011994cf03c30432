#!/usr/bin/env python
"""Headline benchmark: mel-frames/sec of the Lip2Speech inference hot path on LRW-shape clips
(BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the whole hot span of demo.py:84-86 (speaker encoder -> video frontend ->
decoder pre-loop -> 300 autoregressive steps -> postnet) over one batch of B=32 synthetic LRW-shape
clips per GPU (29 frames of 96x96, 19 456 audio samples); 300 mel frames are emitted per clip
(decoder.py:412 always runs max_decoder_steps).  value = N*B*300*K / time.

  value : inputs resident in HBM, CUDA-event time per step, L2 flushed between steps, max over ranks
  e2e   : the same span through the public C-ABI host calls (l2s_infer_host_submit / _wait): pinned host inputs,
          H2D + compute + D2H of every step inside the timed region, consecutive steps double-buffered
  roofline : the persistent decode-loop kernel's algorithmic bytes / its CUDA-event time
  cpu_baseline : the CPU oracle (port of the reference, torch CPU fp32, all host threads) on the same workload

Multi-GPU: clips are independent -> the batch is sharded, B per rank, no data-path collective
("scaling": "weak"); torch.distributed (NCCL) is used only for the barrier and the max-over-ranks.
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

import torch  # noqa: E402

T, H, W, S, STEPS_PER_CLIP = 29, 96, 96, 19456, 300
METRIC = "mel-frames/sec (1-GPU inference) on LRW 29-frame clips"


def algorithmic_decode_bytes(b_gpu: int, t: int = 29, min_t: int = 4, steps: int = 300) -> float:
    """weights streamed once per step + per-clip traffic, fp32."""
    per_clip = 2 * t * 512 * 4 + 2 * min_t * 256 * 4 + 2 * 2 * 2 * 512 * 4 + 512 * 4 + 80 * 4
    return float(steps) * (21_002_572 + b_gpu * per_clip)


def ncu_traffic_bytes(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel from the committed ncu --set full capture
    (profiles/decode_kernel_traffic.json), per launch; None if no capture exists for this batch size."""
    try:
        with open(os.path.join(ROOT, "profiles", "decode_kernel_traffic.json")) as f:
            d = json.load(f)
        return d.get(f"B{batch}", {}).get("dram_bytes")
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_run(batch, reps, warmup=1):
    """Times the CPU oracle (port of the reference modules) on `reps` batches; returns (median s, info)."""
    from lip2speech_b200 import spec, synth
    from oracle import l2s_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    w = spec.seeded_state_dict(spec.full_spec(), 1234)
    spk_w = {k[len("speaker_encoder."):]: v for k, v in w.items() if k.startswith("speaker_encoder.")}
    video, wav, g = synth.video(batch, T, H, W), synth.wav(batch, S), synth.gumbel(batch, T)
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            O.demo_span(w, spk_w, video, wav, g, STEPS_PER_CLIP)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference
    cannot travel to the GPU box) on the host cores, same workload/metric.  Rank 0 only."""
    if rank != 0:
        return
    batch = args.batch
    times = cpu_oracle_run(batch, reps=args.steps, warmup=min(args.warmup, 1))
    total = sum(times)
    value = batch * STEPS_PER_CLIP * len(times) / total
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "mel-frames/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LRW-shape batch={batch} inference, T=29, 96x96, 300 decoder steps (BASELINE configs[2] on host CPU)",
                       "batch_per_step": batch},
            "cpu_baseline": {"value": value, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                             "sample": f"{len(times)} full batches of {batch} clips, oracle/l2s_oracle.py (torch CPU fp32)"},
            "e2e": {"value": value, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="Conv3d stem operands: bf16 (BASELINE configs[2]: 'bf16 frontend + fp32 decoder step') or the 3xTF32 fp32-grade path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.steps > 5:
            args.steps = 5          # bounded CPU sample: keep the whole run within a few minutes
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from lip2speech_b200 import _lib, build, sharding, spec, synth

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    B = args.batch
    prec = _lib.PRECISION_BF16 if args.precision == "bf16" else _lib.PRECISION_FP32
    be = _lib.backend(local_rank)
    be.bind_state_dict(spec.seeded_state_dict(spec.full_spec(), 1234), "", 7)
    # identical synthetic inputs on every rank's shard (seeded per rank)
    video_h = synth.video(B, T, H, W, seed=1234 + rank).pin_memory()
    wav_h = synth.wav(B, S, seed=1234 + rank).pin_memory()
    g_h = synth.gumbel(B, T, seed=1234 + rank).pin_memory()
    video, wav, g = video_h.to(dev), wav_h.to(dev), g_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    be.set_profiling(True)
    for _ in range(max(args.warmup, 3)):
        be.infer(video, wav, g, STEPS_PER_CLIP, prec)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = be.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    spans = {k: [] for k in ("speaker", "video", "preloop", "decode_loop", "postnet")}
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # evict L2 between timed steps (outside the event pair)
        ev[i][0].record()
        be.infer(video, wav, g, STEPS_PER_CLIP, prec)
        ev[i][1].record()
        ev[i][1].synchronize()
        for k in spans:
            spans[k].append(be.span_ms(k))
    barrier()
    launches = be.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sharding.max_over_ranks(sum(step_ms), dev)      # device time, max over ranks

    # ---- e2e: public host-buffer calls, H2D + compute + D2H of EVERY step inside the timed region ---------------------
    # The caller streams batches the way demo.py / evaluate.py loop over a DataLoader: l2s_infer_host_submit / _wait with two
    # staging slots, so the clip copy of step i+1 overlaps the compute of step i; every step's inputs come from pinned host
    # memory and every step's mel / lengths are read back to the host.
    mel_h = [torch.empty(B, 80, STEPS_PER_CLIP).pin_memory() for _ in range(2)]
    len_h = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    be.set_profiling(False)
    for i in range(4):                                       # warm both staging slots (their buffers are allocated on first use)
        be.infer_host_submit(i & 1, video_h, wav_h, g_h, mel_h[i & 1], len_h[i & 1], STEPS_PER_CLIP, prec)
        if i > 0:
            be.infer_host_wait((i - 1) & 1)
    be.infer_host_wait(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        be.infer_host_submit(i & 1, video_h, wav_h, g_h, mel_h[i & 1], len_h[i & 1], STEPS_PER_CLIP, prec)
        if i > 0:
            be.infer_host_wait((i - 1) & 1)                  # step i-1's results are on the host
    be.infer_host_wait((args.steps - 1) & 1)
    torch.cuda.synchronize()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    # latency of one synchronous call (no cross-step overlap), for reference
    t0 = time.perf_counter()
    for _ in range(3):
        be.infer_host(video_h, wav_h, g_h, mel_h[0], len_h[0], STEPS_PER_CLIP, prec)
    sync_call_ms = 1e3 * (time.perf_counter() - t0) / 3
    clocks = sampler.stop()

    if rank == 0:
        frames = world * B * STEPS_PER_CLIP * args.steps
        peaks, peak_src = measured_peaks()
        dec_ms = statistics.mean(spans["decode_loop"])
        alg = algorithmic_decode_bytes(B)
        achieved = alg / (dec_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": frames / (total_ms * 1e-3), "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LRW-shape batch={B}/GPU inference: speaker enc + video frontend + decoder (300 steps) + postnet; "
                                   "T=29, 96x96, S=19456 (BASELINE configs[2], north_star target)",
                       "batch_per_gpu": B, "frames_per_clip": STEPS_PER_CLIP, "precision": ("Conv3d stem: bf16 operands / fp32 accumulate on tcgen05 (BASELINE configs[2] 'bf16 frontend'); " if args.precision == "bf16"
                                     else "Conv3d stem: 3xTF32 on tcgen05; ")
                                    + "all other GEMM-shaped layers on tcgen05 with 3xTF32 error compensation, recurrent step 3xTF32 on mma.sync "
                                      "(fp32 storage everywhere; mel rel err 3e-5 vs reference, features 1e-4 with the bf16 stem)",
                       "l2": "flushed between timed steps (256 MiB memset outside the event pair)", "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": frames / e2e_s, "unit": "mel-frames/s",
                    "h2d_bytes_per_step": int(video_h.numel() * 4 + wav_h.numel() * 4 + g_h.numel() * 4),
                    "d2h_bytes_per_step": int(mel_h[0].numel() * 4 + len_h[0].numel() * 8), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "l2s_infer_host_submit / l2s_infer_host_wait (C ABI, pinned host buffers, two staging slots: step i+1's copy "
                           "overlaps step i's compute)",
                    "synchronous_call_ms": sync_call_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": ("decode3_kernel" if be.debug_flag("dec3") == 1 else "decode_persistent_kernel") + " (300 steps, one launch)", "bound": "hbm", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic_bytes(B),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": dec_ms},
            "stage_ms": {k: statistics.mean(v) for k, v in spans.items()},
            "valid_frames_note": "LRW clips carry 77 real mel frames; the reference always emits 300 (x77/300 for 'valid' frames/s)",
        }
        if not args.no_cpu_baseline:
            reps = 3
            times = cpu_oracle_run(B, reps=reps, warmup=1)
            line["cpu_baseline"] = {"value": B * STEPS_PER_CLIP / statistics.median(times), "unit": "mel-frames/s",
                                    "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"median of {reps} full batches of {B} clips through oracle/l2s_oracle.py (torch CPU fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
