"""One train.py-style step (B clips, T=29, M=77) through the mirror modules, for ncu launch lists / timing.
   python tools/train_step_once.py [B] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, modules, spec, synth
from lip2speech_b200.train_step import ClipAdamW, Loss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T, M = 29, 77
w = {k: v for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items() if not k.startswith("speaker_encoder.")}
net = modules.get_network("train"); net.load_state_dict(w, strict=True); net = net.cuda()
opt = ClipAdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-4, weight_decay=1e-6, max_norm=1.0)
video, spk = synth.video(B, T).cuda(), synth.speaker_embedding(B).cuda()
mels = (synth.mel_like(B, M) * 2 - 5).cuda()
gate = torch.zeros(B, M, device="cuda"); gate[:, -2:] = 1
lens = torch.full((B,), T, dtype=torch.long)
for i in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    out = net(video, None, None, mels, lens, None, lens, 0.5, speaker_embedding=spk)
    ev[1].record()
    loss = sum(Loss()(out, (mels, gate)).values())
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(loss):.4f}  forward {ev[0].elapsed_time(ev[1]):.1f} ms  loss+backward {ev[1].elapsed_time(ev[2]):.1f} ms  "
          f"optimizer {ev[2].elapsed_time(ev[3]):.2f} ms  wall {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
