"""Small end-to-end invocations for compute-sanitizer (memcheck / racecheck): both decode kernels, both stem precisions,
the Decoder.forward flavour, the train-step tail."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.full_spec(), 1234), "", 7)
for B, steps in ((9, 3), (2, 3)):
    for prec in (0, 1):
        mel, lengths = be.infer(synth.video(B, 29).cuda(), synth.wav(B).cuda(), synth.gumbel(B, 29).cuda(), steps=steps, precision=prec)
visual, face = synth.visual_features(10, 29, seed=3)
mels = synth.mel_like(10, 4, seed=3)
be.decoder_forward(visual.cuda(), face[:, 0].cuda(), synth.gumbel(10, 29, seed=3).cuda(), mels.cuda(), torch.tensor([True, False, True, False]))
n = 4099
p, g, m, v, vm = (torch.randn(n + 1, device="cuda")[:n] for _ in range(5))
p, g, m, v, vm = (torch.randn(4100, device="cuda") for _ in range(5))
sq = torch.zeros(1, device="cuda")
be.allreduce_grads(g, 0.5, sq)
be.clip_adamw_step(p, g, m, v.abs(), vm.abs(), sq, 1.0, 1e-3, 0.9, 0.999, 1e-8, 1e-2, 1)
be.loss_fwd_bwd(torch.randn(2, 80, 5).cuda(), torch.randn(2, 80, 5).cuda(), torch.randn(2, 5).cuda(), torch.softmax(torch.randn(8, 501), -1).cuda(),
                torch.randn(2, 80, 5).cuda(), torch.zeros(2, 5).cuda())
torch.cuda.synchronize()
print("sanitize run complete", float(mel.abs().max()))
