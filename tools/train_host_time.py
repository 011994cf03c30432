import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from lip2speech_b200 import _lib, modules, spec, synth
from lip2speech_b200.train_step import ClipAdamW, Loss
B, T, M = 8, 29, 77
w = {k: v for k, v in spec.seeded_state_dict(spec.full_spec(), 1234).items() if not k.startswith("speaker_encoder.")}
net = modules.get_network("train"); net.load_state_dict(w, strict=True); net = net.cuda()
opt = ClipAdamW([{"params": net.decoder.parameters()}, {"params": net.encoder.parameters()}], lr=1e-4, weight_decay=1e-6, max_norm=1.0)
video, spk = synth.video(B, T).cuda(), synth.speaker_embedding(B).cuda()
mels = (synth.mel_like(B, M) * 2 - 5).cuda()
gate = torch.zeros(B, M, device="cuda"); gate[:, -2:] = 1
lens = torch.full((B,), T, dtype=torch.long)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.zero_grad()
    out = net(video, None, None, mels, lens, None, lens, 0.5, speaker_embedding=spk)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t1s = time.perf_counter()
    loss = sum(Loss()(out, (mels, gate)).values())
    loss.backward()
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t2s = time.perf_counter()
    opt.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"step {i}: fwd host-issue {1e3*(t1-t0):.1f} ms, fwd done {1e3*(t1s-t0):.1f}; bwd host-issue {1e3*(t2-t1s):.1f}, bwd done {1e3*(t2s-t1s):.1f}; opt {1e3*(t3-t2s):.1f}", flush=True)
