"""Video frontend: fp32-grade (3xTF32) vs bf16 stem — feature drift against the reference goldens and stage time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth, build
build.build()
w = spec.seeded_state_dict(spec.full_spec(), 1234)
be = _lib.backend(0); be.bind_state_dict(w, "", 7)
golden = torch.load("tests/golden/golden_synthetic.pt", weights_only=True)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
for prec in (0, 1):
    f = be.video_fwd(synth.video(2, 29).cuda(), precision=prec).cpu()
    print("precision", prec, "features vs golden", rel(f, golden["A_video_feat"]))
    f = be.video_fwd(synth.video(1, 5, 88, 88, seed=5).cuda(), precision=prec).cpu()
    print("precision", prec, "features 88x88 vs golden", rel(f, golden["D_video_feat"]))
    mel, lengths = be.infer(synth.video(2, 29).cuda(), synth.wav(2).cuda(), synth.gumbel(2, 29).cuda(), precision=prec)
    print("precision", prec, "full span mel vs golden", rel(mel.cpu(), golden["A_mel"]), torch.equal(lengths.cpu(), golden["A_lengths"]))
v = synth.video(32, 29).cuda()
for prec in (0, 1):
    be.video_fwd(v, precision=prec); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): be.video_fwd(v, precision=prec)
    e1.record(); torch.cuda.synchronize()
    print("precision", prec, "video frontend ms (B=32)", e0.elapsed_time(e1) / 5)
