import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.encoder_spec("encoder."), 1234), "", 1)
v = synth.video(32, 29).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = _lib.PRECISION_BF16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else _lib.PRECISION_FP32
for _ in range(n):
    be.video_fwd(v, prec)
torch.cuda.synchronize()
