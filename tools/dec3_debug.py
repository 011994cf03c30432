"""Stage-pipelined decode kernel (decode3.cuh) vs the CPU oracle, plus its per-role turn timing.  GPU box only.
Prints, never asserts.   python tools/dec3_debug.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth, build
from oracle import l2s_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


build.build()
w = spec.seeded_state_dict(spec.full_spec(), 1234)
b3 = _lib.Backend(0); b3.bind_state_dict(w, "", _lib.PART_DECODER)

for B, steps in ((12, 6), (32, 40), (9, 40)):
    visual, face = synth.visual_features(B, 29, seed=21)
    g = synth.gumbel(B, 29, seed=21)
    m3, l3 = b3.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
    torch.cuda.synchronize()
    ref_mel, ref_len = O.decoder_inference(w, visual, face, g, steps=steps)
    print(f"B={B} steps={steps}: dec3 vs oracle {rel(m3.cpu(), ref_mel):.2e}  lengths ok {torch.equal(l3.cpu(), ref_len)} "
          f"abort={b3.debug_flag('dec_abort')}", flush=True)

PB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
visual, face = synth.visual_features(PB, 29, seed=5)
g = synth.gumbel(PB, 29, seed=5)
b3.set_profiling(True)
for _ in range(3):
    mel, lengths = b3.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
torch.cuda.synchronize()
print(f"B={PB}: decode_loop {b3.span_ms('decode_loop'):.3f} ms  preloop {b3.span_ms('preloop'):.3f}  postnet {b3.span_ms('postnet'):.3f}", flush=True)
t = b3.debug_read("dec.timing3", (148, 6))
# CTA order of pack_decode_program3
names = ["D"] * 64 + ["E"] * 43 + ["A.Q"] * 11 + ["A.CQ"] * 6 + ["A.F"] * 5 + ["Battn"] * 16 + ["Bp2"] * 3
for r in ("D", "E", "A.Q", "A.CQ", "A.F", "Battn", "Bp2"):
    idx = [i for i, n in enumerate(names) if n == r]
    tt = t[idx].mean(0) / 1.965e3 / 300 / 4      # us per turn (4 active turns per step at B=32)
    print(f"   {r:6s} per turn: early(kv wait) {tt[0]:.2f} us  wait-late(query wait) {tt[1]:.2f}  late MMAs {tt[2]:.2f}  reduce+epilogue/attention {tt[3]:.2f}"
          f"  sum {float(tt[:4].sum()):.2f}")
m12, _ = b3.decoder_infer(visual[8:20].cuda(), face[8:20, 0].cuda(), g[32:80].cuda())
print("B=12 sub-batch bit-identical to B=32:", torch.equal(m12, mel[8:20]))
b3.set_profiling(False)
for B in (1, 2, 4, 8, 16, 32, 64):
    visual, face = synth.visual_features(B, 29, seed=6)
    g = synth.gumbel(B, 29, seed=6)
    v, f, gg = visual.cuda(), face[:, 0].cuda(), g.cuda()
    b3.set_profiling(True)
    for _ in range(2):
        b3.decoder_infer(v, f, gg)
    torch.cuda.synchronize()
    print(f"B={B}: decode_loop {b3.span_ms('decode_loop'):.2f} ms", flush=True)
