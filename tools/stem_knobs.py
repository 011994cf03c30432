"""Elimination timing of the bf16 Conv3d stem (needs a -DL2S_DEBUG build: L2S_TC_DEBUG_SKIP bits: 1 epilogue, 2 gather, 4 MMA issue,
32 W TMA).  Results are wrong by construction; prints the video-frontend time at B=32.   python tools/stem_knobs.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = """
import sys, os, torch
sys.path.insert(0, %r)
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.encoder_spec('encoder.'), 1234), '', 1)
v = synth.video(32, 29).cuda()
be.video_fwd(v, _lib.PRECISION_BF16); torch.cuda.synchronize()
be.set_profiling(True)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): be.video_fwd(v, _lib.PRECISION_BF16)
e1.record(); torch.cuda.synchronize()
print('skip', os.environ.get('L2S_TC_DEBUG_SKIP'), 'niter', os.environ.get('L2S_TC_DEBUG_NITER'), 'video B=32 ms', round(e0.elapsed_time(e1) / 5, 3), flush=True)
""" % ROOT
import itertools
cases = [(k, 0) for k in (0, 1, 2, 4, 6, 3, 7, 39)] + [(39, n) for n in (1, 2, 5)] + [(0, n) for n in (1, 2, 5)]
for skip, niter in cases:
    env = dict(os.environ, L2S_TC_DEBUG_SKIP=str(skip), L2S_TC_DEBUG_NITER=str(niter))
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    print((r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip().splitlines()[-1:] if r.returncode else "")
