import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.encoder_spec("encoder."), 1234), "", 1)
v = synth.video(32, 29).cuda()
be.video_fwd(v); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3): be.video_fwd(v)
e1.record(); torch.cuda.synchronize()
print("skip", os.environ.get("L2S_TC_DEBUG_SKIP"), "video B=32 ms", e0.elapsed_time(e1) / 3, flush=True)
