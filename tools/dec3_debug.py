"""Stage-pipelined decode kernel (decode3.cuh) vs the row-partitioned kernel (decode.cuh) vs the CPU oracle, plus timing.
GPU box only.  Prints, never asserts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth, build
from oracle import l2s_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


build.build()
w = spec.seeded_state_dict(spec.full_spec(), 1234)
os.environ["L2S_DEC3"] = "1"
b3 = _lib.Backend(0); b3.bind_state_dict(w, "", _lib.PART_DECODER)
os.environ["L2S_DEC3"] = "0"
b2 = _lib.Backend(0); b2.bind_state_dict(w, "", _lib.PART_DECODER)

for B, steps in ((12, 6), (32, 40), (9, 40)):
    visual, face = synth.visual_features(B, 29, seed=21)
    g = synth.gumbel(B, 29, seed=21)
    m3, l3 = b3.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
    m2, l2 = b2.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
    torch.cuda.synchronize()
    ref_mel, ref_len = O.decoder_inference(w, visual, face, g, steps=steps)
    print(f"B={B} steps={steps}: dec3 vs oracle {rel(m3.cpu(), ref_mel):.2e}  dec2 vs oracle {rel(m2.cpu(), ref_mel):.2e}  "
          f"dec3 vs dec2 {rel(m3, m2):.2e}  lengths3 ok {torch.equal(l3.cpu(), ref_len)} lengths2 ok {torch.equal(l2.cpu(), ref_len)}", flush=True)

import sys as _sys
PB = int(_sys.argv[1]) if len(_sys.argv) > 1 else 32
visual, face = synth.visual_features(PB, 29, seed=5)
g = synth.gumbel(PB, 29, seed=5)
for name, be in (("dec3", b3), ("dec2", b2)):
    be.set_profiling(True)
    for _ in range(3):
        mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
    torch.cuda.synchronize()
    print(f"{name}: decode_loop {be.span_ms('decode_loop'):.3f} ms  preloop {be.span_ms('preloop'):.3f}  postnet {be.span_ms('postnet'):.3f}", flush=True)
    if name == "dec3":
        try:
            t = be.debug_read("dec.timing3", (148, 6))
            role = ["B"] * 148
            # CTA order of pack_decode_program3
            names = ["D"] * 64 + ["E"] * 43 + ["A.Q"] * 11 + ["A.CQ"] * 6 + ["A.F"] * 5 + ["Battn"] * 16 + ["Bp2"] * 3
            for r in ("D", "E", "A.Q", "A.CQ", "A.F", "Battn", "Bp2"):
                idx = [i for i, n in enumerate(names) if n == r]
                tt = t[idx].mean(0) / 1.965e3 / 300      # us per step (4 turns)
                print(f"   {r:6s} per turn: early {tt[0]/4:.2f} us  wait {tt[1]/4:.2f}  late-acc {tt[2]/4:.2f}  reduce+epi/attn {tt[3]/4:.2f}  arrive {tt[5]/4:.2f}  (idle turns {tt[4]/4:.2f})")
        except Exception as e:
            print("   timing read failed:", e)
mel3, _ = b3.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
mel2, _ = b2.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
print("B=32, 300 steps: dec3 vs dec2", rel(mel3, mel2))
m12, _ = b3.decoder_infer(visual[8:20].cuda(), face[8:20, 0].cuda(), g[32:80].cuda())
print("B=12 sub-batch bit-identical to B=32:", torch.equal(m12, mel3[8:20]))

for B in (1, 2, 4, 8, 64):
    visual, face = synth.visual_features(B, 29, seed=6)
    g = synth.gumbel(B, 29, seed=6)
    v, f, gg = visual.cuda(), face[:, 0].cuda(), g.cuda()
    line = f"B={B}:"
    for name, be in (("dec3", b3), ("dec2", b2)):
        be.set_profiling(True)
        for _ in range(2):
            mel, _ = be.decoder_infer(v, f, gg)
        torch.cuda.synchronize()
        line += f"  {name} decode_loop {be.span_ms('decode_loop'):.2f} ms"
    print(line, flush=True)
