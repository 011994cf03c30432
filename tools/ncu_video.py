import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.encoder_spec("encoder."), 1234), "", 1)
v = synth.video(32, 29).cuda()
for _ in range(2):
    be.video_fwd(v)
torch.cuda.synchronize()
