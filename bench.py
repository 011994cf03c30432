#!/usr/bin/env python
"""Headline benchmark: mel-frames/sec of the Lip2Speech inference hot path on LRW-shape clips
(BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c4|train-tail] [--batch B]

One "step" = one pass of the whole hot span of demo.py:84-86 (speaker encoder -> video frontend ->
decoder pre-loop -> 300 autoregressive steps -> postnet) over one batch of synthetic clips per GPU;
300 mel frames are emitted per clip (decoder.py:412 always runs max_decoder_steps).  value = clips*300*K / time.

  --config c2 (default) : BASELINE configs[2] — B=32 clips per GPU, T=29 frames of 96x96, S=19 456 samples (weak scaling)
  --config c1           : BASELINE configs[1] — single-clip latency, B=1
  --config c4           : BASELINE configs[4] — AVSpeech shape, T=75, S=48 000, B=256 TOTAL split over the GPUs (strong scaling)
  --config c3           : BASELINE configs[3] — train.py step (train.py:167-193): train-mode forward (video + decoder, M=77 teacher
                          frames) + Loss + backward + gradient all-reduce + clip + AdamW, batch 64 TOTAL split over the GPUs
                          (8 clips per GPU at N=8; at N=1 all 64 on one GPU in 8-clip micro-batches is NOT done: one GPU runs 8)
  --config train-tail   : the train step's flat-buffer tail (train.py:184-193): gradient all-reduce of 38.44 M fp32 over the
                          ranks (NCCL, the path's only collective) + 1/world + norm + clip + AdamW(amsgrad); its own metric (ms)

  value : inputs resident in HBM, CUDA-event time per step, L2 flushed between steps, max over ranks
  e2e   : the same span through the public C-ABI host calls (l2s_infer_host_submit_u8 / _wait): pinned host inputs — raw uint8
          frames as datasets/lrw/dataset.py:20-24 decodes them, /255 + Normalize fused on the device — H2D + compute + D2H of
          every step inside the timed region, consecutive steps double-buffered (e2e_f32: the fp32 NCDHW host entry point)
  roofline : the persistent decode-loop kernel's algorithmic bytes / its CUDA-event time
  cpu_baseline : the CPU oracle (port of the reference, torch CPU fp32, all host threads) on a bounded sample of the workload
  gpu_eager_baseline : the same oracle port on .cuda() tensors (PyTorch eager: cuDNN / cuBLAS kernels) — the incumbent GPU path

Multi-GPU: clips are independent -> the batch is sharded, no data-path collective; torch.distributed (NCCL) is used only for
the barrier and the max-over-ranks (train-tail: the gradient all-reduce goes through the library's own communicator).
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

H, W, STEPS_PER_CLIP = 96, 96, 300
METRIC = "mel-frames/sec (1-GPU inference) on LRW 29-frame clips"
CONFIGS = {
    "c1": dict(T=29, S=19456, per_gpu=1, total=None, scaling="weak",
               name="LRW single-clip inference, batch=1 (BASELINE configs[1])"),
    "c2": dict(T=29, S=19456, per_gpu=32, total=None, scaling="weak",
               name="LRW-shape batch=32/GPU inference (BASELINE configs[2], north_star target)"),
    "c4": dict(T=75, S=48000, per_gpu=None, total=256, scaling="strong",
               name="AVSpeech-shape 75-frame clips, batch=256 total sharded over the GPUs (BASELINE configs[4])"),
}
N_TRAIN_PARAMS = 38_436_836          # decoder 37 285 512 + video 1 151 324


def min_t_of(t: int) -> int:
    return min([t] + [(t - k) // k + 1 for k in (1, 3, 5, 7)])


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
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "FALLBACK (B200_PROFILING.md; MEASURED_PEAKS.json missing)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, rank, world):
        """One poller for the whole job, on rank 0, over the GPUs of all local ranks: every NVML query takes a driver-wide lock, and
        eight pollers at 10 Hz next to eight ranks that launch a graph every 25 ms showed up as rank-to-rank jitter at N = 8."""
        self.index, self.lines, self.proc = ",".join(str(i) for i in range(world)), [], None
        self.enabled = rank == 0

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", self.index], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def oracle_run(batch, t, s_len, reps, warmup=1, device="cpu"):
    """Times the oracle (port of the reference modules) on `reps` batches; returns the list of seconds per batch."""
    from lip2speech_b200 import spec, synth
    from oracle import l2s_oracle as O
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    w = {k: v.to(device) for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items()}
    spk_w = {k[len("speaker_encoder."):]: v for k, v in w.items() if k.startswith("speaker_encoder.")}
    video, wav, g = (x.to(device) for x in (synth.video(batch, t, H, W), synth.wav(batch, s_len), synth.gumbel(batch, t)))
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            O.demo_span(w, spk_w, video, wav, g, STEPS_PER_CLIP)
            if device != "cpu":
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def workload(args, world):
    cfg = CONFIGS[args.config]
    if args.batch:
        b_gpu = args.batch
    elif cfg["total"]:
        if cfg["total"] % world:
            raise SystemExit(f"--config {args.config}: {cfg['total']} clips do not split over {world} GPUs")
        b_gpu = cfg["total"] // world
    else:
        b_gpu = cfg["per_gpu"]
    return cfg, b_gpu


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference
    cannot travel to the GPU box) on the host cores, same workload/metric.  Rank 0 only."""
    if rank != 0:
        return
    cfg, b_gpu = workload(args, world)
    batch = min(b_gpu, 32)            # bounded sample of the workload: at most 32 clips per timed batch
    times = oracle_run(batch, cfg["T"], cfg["S"], reps=args.steps, warmup=min(args.warmup, 1))
    total = sum(times)
    value = batch * STEPS_PER_CLIP * len(times) / total
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "mel-frames/s", "n_gpus": 0, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}: T={cfg['T']}, 96x96, S={cfg['S']}, 300 decoder steps — on the host CPU, {batch} clips per timed batch",
                       "batch_per_step": batch, "gpus_requested": args.gpus},
            "cpu_baseline": {"value": value, "unit": "mel-frames/s", "cores": cores, "kind": "port",
                             "sample": f"{len(times)} batches of {batch} clips, oracle/l2s_oracle.py (torch CPU fp32)"},
            "e2e": {"value": value, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def oracle_train_step(batch, T, M, reps, device="cpu"):
    """Baseline arms of --config c3: autograd through the train-mode oracle (port of the reference's train step, torch fp32) +
    clip_grad_norm_ + torch.optim.AdamW(amsgrad) — on the host cores (device="cpu") or, as the incumbent GPU path, with every tensor
    on the GPU (PyTorch eager: cuDNN / cuBLAS kernels, one launch per op)."""
    from lip2speech_b200 import spec, synth
    from oracle import train_oracle as TO
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device(device)
    w = {k: v for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items() if not k.startswith("speaker_encoder.")}
    sd = {k: (v.clone().to(dev).requires_grad_(True) if v.is_floating_point() and not spec.is_buffer(k) else v.clone().to(dev)) for k, v in w.items()}
    params = [p for p in sd.values() if torch.is_tensor(p) and p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-6, amsgrad=True)
    video, spk = synth.video(batch, T, H, W).to(dev), synth.speaker_embedding(batch).to(dev)
    mels = (synth.mel_like(batch, M) * 2 - 5).to(dev)
    gate_t = torch.zeros(batch, M); gate_t[:, -2:] = 1
    gate_t = gate_t.to(dev)
    times = []
    for i in range(reps + 1):
        noise = TO.reference_noise(batch, T, M, 0.5, with_video=True, generator=torch.Generator().manual_seed(i)).to(dev)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        out = TO.lip2speech_forward_train(sd, video, spk, mels, noise)
        sum(TO.loss_forward(out, (mels, gate_t)).values()).backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if i > 0:
            times.append(time.perf_counter() - t0)
    return times


def run_train_step(args, rank, local_rank, world):
    """--config c3: one train.py iteration per step through the reference-shaped mirror modules: net(...) in train() ->
    Loss -> loss.backward() -> ClipAdamW.step() (gradient all-reduce over the ranks inside).  value = clips x M teacher
    frames per second over all ranks; inputs resident in HBM; e2e = the same with the batch coming from pinned host memory and
    the loss read back every step."""
    import torch.distributed as dist
    from lip2speech_b200 import _lib, modules, sharding, spec, synth
    from lip2speech_b200.train_step import ClipAdamW, Loss, init_data_parallel
    dev = torch.device("cuda", local_rank)
    total = 64
    B = args.batch or max(1, total // max(world, 8))          # 8 clips per GPU (64 over 8 GPUs); fewer GPUs keep 8 per GPU
    T, M = 29, 77
    be = _lib.backend(local_rank)
    if world > 1:
        init_data_parallel(be, rank, world)
    w = {k: v for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items() if not k.startswith("speaker_encoder.")}
    net = modules.get_network("train")
    net.load_state_dict(w, strict=True)
    net = net.to(dev)
    opt = ClipAdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-4, weight_decay=1e-6, max_norm=1.0, backend=be)
    loss_fn = Loss()
    video_h = synth.video(B, T, H, W, seed=1234 + rank).pin_memory()
    spk_h = synth.speaker_embedding(B, seed=1234 + rank).pin_memory()
    mels_h = (synth.mel_like(B, M, seed=1234 + rank) * 2 - 5).pin_memory()
    gate_h = torch.zeros(B, M); gate_h[:, -2:] = 1
    gate_h = gate_h.pin_memory()
    video, spk, mels, gate = video_h.to(dev), spk_h.to(dev), mels_h.to(dev), gate_h.to(dev)
    lens = torch.full((B,), T, dtype=torch.long)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(v, s_, m_, g_):
        opt.zero_grad()
        out = net(v, None, None, m_, lens, None, lens, 0.5, speaker_embedding=s_)
        loss = sum(loss_fn(out, (m_, g_)).values())
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step(video, spk, mels, gate)
    barrier()
    sampler = ClockSampler(rank, world); sampler.start()
    launches0 = be.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        step(video, spk, mels, gate)
        ev[i][1].record()
    barrier()
    launches = be.launch_count() - launches0
    total_ms = sharding.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev)
    # end to end: every step's batch comes from pinned host memory and every step's loss is read on the host — one step late
    # (the copy of step i's loss is awaited after step i+1 has been issued, as a training loop that logs asynchronously does), so the
    # host's ~3 ms of Python per step overlap the device's work instead of serialising with it
    loss_h = [torch.zeros(1).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    seen = []
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(video_h.to(dev, non_blocking=True), spk_h.to(dev, non_blocking=True), mels_h.to(dev, non_blocking=True), gate_h.to(dev, non_blocking=True))
        loss_h[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        done[i & 1].record()
        if i > 0:
            done[(i - 1) & 1].synchronize()
            seen.append(float(loss_h[(i - 1) & 1]))
    done[(args.steps - 1) & 1].synchronize()
    seen.append(float(loss_h[(args.steps - 1) & 1]))
    torch.cuda.synchronize()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    assert len(seen) == args.steps and all(v == v for v in seen), "every step's loss must have reached the host"
    clocks = sampler.stop()
    if rank == 0:
        frames = world * B * M * args.steps
        line = {"metric": "mel-frames/sec (train step: forward + backward + all-reduce + clip + AdamW) on LRW 29-frame clips, M=77 teacher frames",
                "value": frames / (total_ms * 1e-3), "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (fp32 FMA kernels; the decoder's large GEMMs as 3xTF32 tensor-core products with fp32 accumulation)", "data": "synthetic",
                "config": {"workload": f"train.py:167-193 step, {B} clips per GPU (BASELINE configs[3]: batch 64 over 8 GPUs), T=29, 96x96, M=77; "
                                       "train-mode video frontend + decoder (BatchNorm batch statistics, all dropout sites, BPTT), Loss, "
                                       "gradient all-reduce, clip_grad_norm_(1.0), AdamW(amsgrad)", "name": "c3", "batch_per_gpu": B,
                           "l2": "flushed between timed steps", "parallelism": f"data-parallel x{world}, one ncclAllReduce of 153.7 MB per step"},
                "e2e": {"value": frames / e2e_s, "unit": "mel-frames/s", "ms_per_step": 1e3 * e2e_s / args.steps,
                        "h2d_bytes_per_step": int(4 * (video_h.numel() + spk_h.numel() + mels_h.numel() + gate_h.numel())), "d2h_bytes_per_step": 4,
                        "api": "mirror modules (modules.Lip2Speech in train()) + train_step.Loss + ClipAdamW: pinned host batch in every step, every step's loss read on the host (awaited one step late)"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"kernel": "sgemm_kernel (SIMT fp32 GEMM of the train path)", "bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s",
                             "frac": None, "traffic": None,
                             "note": "fp32 train kernels replayed as CUDA graphs: a chain of ~5 K dependent small launches per step (profiles/r2_train_step_kernel_shares.txt); no single kernel dominates, so no roofline claim is made for this config"}}
        if not args.no_eager_baseline:
            times = oracle_train_step(B, T, M, reps=3, device=f"cuda:{local_rank}")
            line["gpu_eager_baseline"] = {"value": B * M / statistics.median(times), "unit": "mel-frames/s", "ms_per_step": 1e3 * statistics.median(times),
                                          "sample": f"3 train steps of {B} clips: the same oracle port with every tensor on the GPU (PyTorch eager autograd, "
                                                    "cuDNN / cuBLAS kernels) + clip_grad_norm_ + torch.optim.AdamW(amsgrad) — the incumbent GPU path"}
        if not args.no_cpu_baseline:
            times = oracle_train_step(min(B, 4), T, M, reps=1)
            line["cpu_baseline"] = {"value": min(B, 4) * M / statistics.median(times), "unit": "mel-frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"1 train step of {min(B, 4)} clips: autograd through oracle/train_oracle.py (torch CPU fp32) + clip + AdamW"}
        print(json.dumps(line), flush=True)
    if world > 1:
        be.comm_destroy()


def run_train_tail(args, rank, local_rank, world):
    """The one collective of the path: between loss.backward() (train.py:184) and clip_grad_norm_/optim.step() (191-193) the
    flat 38.44 M-float gradient is summed over the ranks (ncclAllReduce through the library's communicator), scaled by
    1/world with its squared norm produced in the same pass, then clip + AdamW(amsgrad) in one streaming pass."""
    import torch.distributed as dist
    from lip2speech_b200 import _lib, sharding
    from lip2speech_b200.train_step import init_data_parallel
    dev = torch.device("cuda", local_rank)
    be = _lib.backend(local_rank)
    if world > 1:
        init_data_parallel(be, rank, world)
    n = (N_TRAIN_PARAMS + 3) // 4 * 4
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    p = torch.randn(n, device=dev, generator=gen) * 0.05
    g0 = torch.randn(n, device=dev, generator=gen) * 1e-3
    g = g0.clone()
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    sq = torch.zeros(1, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(t):
        be.allreduce_grads(g, 1.0 / world, sq)
        be.clip_adamw_step(p, g, m, v, vmax, sq, 1.0, 1e-4, 0.9, 0.999, 1e-8, 1e-6, t)

    for t in range(max(args.warmup, 3)):
        g.copy_(g0); step(t + 1)
    barrier()
    sampler = ClockSampler(rank, world); sampler.start()
    launches0 = be.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for i in range(args.steps):
        g.copy_(g0)
        flush.zero_()
        if world > 1:
            dist.barrier()                  # ranks enter the collective together, as after a backward pass of equal length
        ev[i][0].record()
        be.allreduce_grads(g, 1.0 / world, sq)
        ev[i][1].record()
        be.clip_adamw_step(p, g, m, v, vmax, sq, 1.0, 1e-4, 0.9, 0.999, 1e-8, 1e-6, 100 + i)
        ev[i][2].record()
    barrier()
    launches = be.launch_count() - launches0
    ar = sharding.max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in ev), dev) / args.steps
    up = sharding.max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in ev), dev) / args.steps
    tot = sharding.max_over_ranks(sum(e[0].elapsed_time(e[2]) for e in ev), dev) / args.steps
    # e2e: the public call sequence of the host mirror (ClipAdamW.step) with the step's scalar result (the gradient norm,
    # what clip_grad_norm_ returns to train.py:191) read back to the host every step
    t0 = time.perf_counter()
    for i in range(args.steps):
        g.copy_(g0)
        step(200 + i)
        float(sq.sqrt())
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    clocks = sampler.stop()
    if rank == 0:
        peaks, peak_src = measured_peaks()
        nbytes = n * 4
        bus = (2 * (world - 1) / world * nbytes / (ar * 1e-3) / 1e9) if world > 1 else None
        line = {"metric": "train-step tail: gradient all-reduce (153.7 MB fp32) + 1/world + norm + clip + AdamW(amsgrad), ms per step",
                "value": tot, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": tot,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "train.py:184-193 tail on the flat 38 436 836-parameter buffer (decoder + video), one rank per GPU",
                           "l2": "flushed between timed steps", "collective": "ncclAllReduce(sum) over NVLink/NVSwitch via l2s_allreduce_grads"},
                "collective": {"allreduce_scale_norm_ms": ar, "bus_gbs": bus, "bytes": nbytes, "world": world},
                "update_ms": up,
                "roofline": {"kernel": "clip_adamw_kernel (40 B per parameter: g, p, m, v, vmax read + written)", "bound": "hbm",
                             "achieved": 40 * n / (up * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None,
                             "peak_source": peak_src},
                "e2e": {"value": 1e3 * e2e_s / args.steps, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                        "api": "l2s_allreduce_grads + l2s_clip_adamw_step + gradient-norm readback (ClipAdamW.step)"},
                "gpu_launches": int(launches), "clocks": clocks}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / peaks["hbm_gbs"]
        print(json.dumps(line), flush=True)
    if world > 1:
        be.comm_destroy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["c3", "train-tail"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU per step (default: the config's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="Conv3d stem operands: bf16 (BASELINE configs[2]: 'bf16 frontend + fp32 decoder step') or the 3xTF32 fp32-grade path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.config == "train-tail":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference has no distributed code (SURVEY.md §2.1): the train-tail config has no reference arm"}))
            return
        if args.config == "c3":
            if rank == 0:
                b = min(args.batch or 8, 4)
                times = oracle_train_step(b, 29, 77, reps=max(1, min(args.steps, 2)))
                v = b * 77 * len(times) / sum(times)
                print(json.dumps({"impl": "reference", "metric": "mel-frames/sec (train step) on LRW 29-frame clips, M=77", "value": v, "unit": "mel-frames/s",
                                  "n_gpus": 0, "steps": len(times), "warmup": 1, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
                                  "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                                  "config": {"workload": f"train.py:167-193 step on the host CPU, {b} clips per timed step (autograd through the oracle port)", "name": "c3"},
                                  "cpu_baseline": {"value": v, "unit": "mel-frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                                   "sample": f"{len(times)} train steps of {b} clips"},
                                  "e2e": {"value": v, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
            return
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
    if args.config in ("train-tail", "c3"):
        (run_train_tail if args.config == "train-tail" else run_train_step)(args, rank, local_rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    cfg, B = workload(args, world)
    T, S = cfg["T"], cfg["S"]
    min_t = min_t_of(T)
    prec = _lib.PRECISION_BF16 if args.precision == "bf16" else _lib.PRECISION_FP32
    be = _lib.backend(local_rank)
    be.bind_state_dict(spec.seeded_state_dict(spec.full_spec(), 1234), "", 7)
    # synthetic inputs, seeded per rank: raw uint8 frames (what the dataset decodes) and their normalised fp32 form
    frames_h = synth.frames_u8(B, T, H, W, seed=1234 + rank).pin_memory()
    wav_h = synth.wav(B, S, seed=1234 + rank).pin_memory()
    g_h = synth.gumbel(B, T, seed=1234 + rank).pin_memory()
    video = synth.normalise_frames(frames_h).to(dev)
    wav, g = wav_h.to(dev), g_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    be.set_profiling(True)
    for _ in range(max(args.warmup, 3)):
        be.infer(video, wav, g, STEPS_PER_CLIP, prec)
    barrier()

    sampler = ClockSampler(rank, world)
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
    # The caller streams batches the way demo.py / evaluate.py loop over a DataLoader: l2s_infer_host_submit_u8 / _wait with
    # two staging slots, so the clip copy of step i+1 overlaps the compute of step i; every step's inputs come from pinned host
    # memory and every step's mel / lengths are read back to the host.
    mel_h = [torch.empty(B, 80, STEPS_PER_CLIP).pin_memory() for _ in range(2)]
    len_h = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
    be.set_profiling(False)

    def e2e_loop(submit, first_arg):
        for i in range(4):                                   # warm both staging slots (their buffers are allocated on first use)
            submit(i & 1, first_arg, wav_h, g_h, mel_h[i & 1], len_h[i & 1], STEPS_PER_CLIP, prec)
            if i > 0:
                be.infer_host_wait((i - 1) & 1)
        be.infer_host_wait(1)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            submit(i & 1, first_arg, wav_h, g_h, mel_h[i & 1], len_h[i & 1], STEPS_PER_CLIP, prec)
            if i > 0:
                be.infer_host_wait((i - 1) & 1)              # step i-1's results are on the host
        be.infer_host_wait((args.steps - 1) & 1)
        torch.cuda.synchronize()
        return sharding.max_over_ranks(time.perf_counter() - t0, dev)

    e2e_s = e2e_loop(be.infer_host_submit_u8, frames_h)
    # latency of one synchronous call (no cross-step overlap), for reference
    t0 = time.perf_counter()
    for _ in range(3):
        be.infer_host_submit_u8(0, frames_h, wav_h, g_h, mel_h[0], len_h[0], STEPS_PER_CLIP, prec)
        be.infer_host_wait(0)
    sync_call_ms = 1e3 * (time.perf_counter() - t0) / 3
    e2e_f32_s = None
    if B * T <= 32 * 29:                                     # the fp32 NCDHW host entry point (4x the clip bytes), small configs only
        video_h = video.cpu().pin_memory()
        e2e_f32_s = e2e_loop(be.infer_host_submit, video_h)
    clocks = sampler.stop()

    if rank == 0:
        frames = world * B * STEPS_PER_CLIP * args.steps
        peaks, peak_src = measured_peaks()
        dec_ms = statistics.mean(spans["decode_loop"])
        nchunks = -(-B // 32)                                # decode launches per pass (32 clips each)
        alg = sum(algorithmic_decode_bytes(min(32, B - 32 * j), T, min_t) for j in range(nchunks))
        achieved = alg / (dec_ms * 1e-3) / 1e9
        stem_note = ("Conv3d stem: bf16 operands / fp32 accumulate on tcgen05 (BASELINE configs[2] 'bf16 frontend'); " if args.precision == "bf16"
                     else "Conv3d stem: 3xTF32 on tcgen05; ")
        line = {
            "metric": METRIC, "value": frames / (total_ms * 1e-3), "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "bf16 stem + f32 (3xTF32) elsewhere" if args.precision == "bf16" else "f32 (3xTF32)", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}: speaker enc + video frontend + decoder (300 steps) + postnet; T={T}, 96x96, S={S}, {B} clips per GPU",
                       "name": args.config, "batch_per_gpu": B, "frames_per_clip": STEPS_PER_CLIP,
                       "precision": stem_note + "all other GEMM-shaped layers on tcgen05 with 3xTF32 error compensation, recurrent step 3xTF32 on "
                                                "mma.sync (fp32 storage everywhere; mel rel err 3e-5 vs reference, features 1e-4 with the bf16 stem)",
                       "l2": "flushed between timed steps (256 MiB memset outside the event pair)", "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": frames / e2e_s, "unit": "mel-frames/s",
                    "h2d_bytes_per_step": int(frames_h.numel() + wav_h.numel() * 4 + g_h.numel() * 4),
                    "d2h_bytes_per_step": int(mel_h[0].numel() * 4 + len_h[0].numel() * 8), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "l2s_infer_host_submit_u8 / l2s_infer_host_wait (C ABI, pinned host buffers: raw uint8 frames [B,T,H,W,3] as the dataset "
                           "decodes them, /255 + Normalize fused on the device; two staging slots: step i+1's copy overlaps step i's compute)",
                    "synchronous_call_ms": sync_call_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": f"decode3_kernel (300 steps, {nchunks} launch{'es' if nchunks > 1 else ''} of <= 32 clips)", "bound": "hbm", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic_bytes(B),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg / nchunks, "kernel_ms": dec_ms / nchunks,
                         "launches_per_step": nchunks},
            "stage_ms": {k: statistics.mean(v) for k, v in spans.items()},
            "ms_per_clip": total_ms / args.steps / B,
            "valid_frames_note": "LRW clips carry 77 real mel frames; the reference always emits 300 (x77/300 for 'valid' frames/s)",
        }
        if e2e_f32_s is not None:
            line["e2e_f32"] = {"value": frames / e2e_f32_s, "unit": "mel-frames/s", "ms_per_step": 1e3 * e2e_f32_s / args.steps,
                               "h2d_bytes_per_step": int(video.numel() * 4 + wav_h.numel() * 4 + g_h.numel() * 4),
                               "api": "l2s_infer_host_submit / _wait (normalised fp32 NCDHW clips from the host, the r1 arm)"}
        sample_b = min(B, 32)
        if not args.no_eager_baseline:
            del flush
            torch.cuda.empty_cache()
            reps = 3
            times = oracle_run(sample_b, T, S, reps=reps, warmup=1, device=dev)
            line["gpu_eager_baseline"] = {"value": sample_b * STEPS_PER_CLIP / statistics.median(times), "unit": "mel-frames/s",
                                          "ms_per_batch": 1e3 * statistics.median(times), "kind": "port on CUDA (PyTorch eager: cuDNN / cuBLAS / ATen kernels)",
                                          "sample": f"median of {reps} batches of {sample_b} clips through oracle/l2s_oracle.py on the same GPU "
                                                    "(it has no per-step host sync, unlike decoder.py:430 — a favourable reading of the incumbent)"}
        if not args.no_cpu_baseline:
            reps = 3
            times = oracle_run(sample_b, T, S, reps=reps, warmup=1)
            line["cpu_baseline"] = {"value": sample_b * STEPS_PER_CLIP / statistics.median(times), "unit": "mel-frames/s",
                                    "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"median of {reps} batches of {sample_b} clips through oracle/l2s_oracle.py (torch CPU fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
