"""Stage-by-stage parity report (GPU box).  Prints rel errors for every intermediate; never asserts."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth, build
from oracle import l2s_oracle as O

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

def step(name, fn):
    try:
        t = time.time(); fn(); torch.cuda.synchronize(); print(f"[ok] {name} ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[FAIL] {name}: {e}", flush=True); traceback.print_exc()

build.build()
w = spec.seeded_state_dict(spec.full_spec(), 1234)
spk_w = {k[len("speaker_encoder."):]: v for k, v in w.items() if k.startswith("speaker_encoder.")}
be = _lib.backend(0)
t = time.time(); be.bind_state_dict(w, "", 7); print("bind+commit", time.time() - t, flush=True)
golden = torch.load("tests/golden/golden_synthetic.pt", weights_only=True)

def spk():
    wav = synth.wav(2)
    raw = be.speaker_fwd(wav.cuda(), False).cpu()
    print("  speaker raw vs golden", rel(raw, golden["A_spk_raw"]))
step("speaker", spk)

def vid():
    f = be.video_fwd(synth.video(2, 29).cuda()).cpu()
    print("  video feat vs golden", rel(f, golden["A_video_feat"]))
    f = be.video_fwd(synth.video(1, 5, 88, 88, seed=5).cuda()).cpu()
    print("  video feat 88 vs golden", rel(f, golden["D_video_feat"]))
step("video", vid)

def post():
    y = be.postnet_fwd(synth.mel_like(2, 77).cuda()).cpu()
    print("  postnet vs golden", rel(y, golden["E_postnet"]))
step("postnet", post)

def dec():
    visual, face = synth.visual_features(3, 29); g = synth.gumbel(3, 29)
    pre = O.decoder_preloop(w, visual, face[:, 0], g)
    for steps in (1, 2, 5, 300):
        mel, lengths = be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda(), steps=steps)
        torch.cuda.synchronize()
        if steps == 1:
            for nm, shape, ref in (("dec.enc", (3, 29, 512), pre["enc"]), ("dec.enc_cell", (3, 512), pre["enc_cell"]),
                                   ("dec.K", (3, 29, 512), pre["k"].permute(0, 2, 1)), ("dec.V", (3, 29, 512), pre["v"]),
                                   ("dec.ckey", (3, 4, 256), pre["ckey"].permute(0, 2, 1)), ("dec.cval", (3, 4, 256), pre["cval"])):
                print("  ", nm, rel(be.debug_read(nm, shape), ref))
        outs, lens, _ = O.decoder_steps(w, pre, steps)
        print(f"  steps={steps} outputs", rel(be.debug_read("dec.outputs", (3, steps, 80)), outs), "lengths", lengths.tolist(), lens.tolist())
    print("  mel vs golden", rel(mel.cpu(), golden["B_mel"]))
step("decoder", dec)

def full():
    mel, lengths = be.infer(synth.video(2, 29).cuda(), synth.wav(2).cuda(), synth.gumbel(2, 29).cuda())
    print("  full span vs golden", rel(mel.cpu(), golden["A_mel"]), lengths.tolist())
step("full", full)

def timing():
    for B in (1, 32):
        video, wav, g = synth.video(B, 29).cuda(), synth.wav(B).cuda(), synth.gumbel(B, 29).cuda()
        visual, face = synth.visual_features(B, 29); visual, spk_e, = visual.cuda(), face[:, 0].cuda()
        def tm(fn, n=3):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        print(f"  B={B}: speaker {tm(lambda: be.speaker_fwd(wav, True)):.3f} ms, video {tm(lambda: be.video_fwd(video)):.3f} ms, "
              f"decoder {tm(lambda: be.decoder_infer(visual, spk_e, g)):.3f} ms, full {tm(lambda: be.infer(video, wav, g)):.3f} ms", flush=True)
step("timing", timing)

def stage_timing():
    be.set_profiling(True)
    for B in (1, 32):
        visual, face = synth.visual_features(B, 29)
        g = synth.gumbel(B, 29)
        be.decoder_infer(visual.cuda(), face[:, 0].cuda(), g.cuda())
        torch.cuda.synchronize()
        tm = be.debug_read("dec.timing", (148, 12)) / 1965.0 / 300.0     # us per step
        names = ["B early", "B wait", "B late", "D early", "D wait", "D late", "E early", "E wait", "E late", "A early", "A wait", "A late"]
        print(f"  B={B}: decode_loop {be.span_ms('decode_loop'):.3f} ms; per-step us (mean / max over CTAs):")
        for i, n in enumerate(names):
            print(f"     {n:7s} {tm[:, i].mean():7.3f} {tm[:, i].max():7.3f}   cta0={tm[0, i]:.3f} cta147={tm[147, i]:.3f}")
        print("     sum(mean) =", float(tm.mean(0).sum()))
    be.set_profiling(False)
step("stage_timing", stage_timing)
