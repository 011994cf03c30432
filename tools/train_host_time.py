"""Host-side (Python + launch) time of one train step against its device time: the step is GPU-bound only while the host
finishes issuing step i+1 before the GPU finishes step i.   python tools/train_host_time.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, modules, spec, synth
from lip2speech_b200.train_step import ClipAdamW, Loss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T, M = 29, 77
w = {k: v for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items() if not k.startswith("speaker_encoder.")}
net = modules.get_network("train"); net.load_state_dict(w, strict=True); net = net.cuda()
opt = ClipAdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-4, weight_decay=1e-6, max_norm=1.0)
video, spk = synth.video(B, T).cuda(), synth.speaker_embedding(B).cuda()
mels = (synth.mel_like(B, M) * 2 - 5).cuda()
gate = torch.zeros(B, M, device="cuda"); gate[:, -2:] = 1
lens = torch.full((B,), T, dtype=torch.long)
loss_fn = Loss()
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    t1 = time.perf_counter()
    out = net(video, None, None, mels, lens, None, lens, 0.5, speaker_embedding=spk)
    t2 = time.perf_counter()
    loss = sum(loss_fn(out, (mels, gate)).values())
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    opt.step()
    t5 = time.perf_counter()
    torch.cuda.synchronize(); t6 = time.perf_counter()
    print(f"step {i}: host zero_grad {1e3*(t1-t0):.2f}  forward {1e3*(t2-t1):.2f}  loss {1e3*(t3-t2):.2f}  backward {1e3*(t4-t3):.2f}  optimizer {1e3*(t5-t4):.2f}"
          f"  = host {1e3*(t5-t0):.2f} ms; device done after {1e3*(t6-t0):.2f} ms", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    opt.zero_grad()
    out = net(video, None, None, mels, lens, None, lens, 0.5, speaker_embedding=spk)
    loss = sum(loss_fn(out, (mels, gate)).values())
    loss.backward()
    opt.step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
