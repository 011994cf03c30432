"""Minimal driver for ncu: one decoder_infer at B=32 (the decode_persistent_kernel is the target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.decoder_spec("decoder."), 1234), "", 4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
visual, face = synth.visual_features(B, 29)
g = synth.gumbel(B, 29)
for _ in range(2):
    be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
torch.cuda.synchronize()
