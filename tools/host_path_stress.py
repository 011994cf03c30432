"""Stress the host-buffer entry points (sync call and two-slot submit/wait) against the device-pointer path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, spec, synth
be = _lib.backend(0)
be.bind_state_dict(spec.seeded_state_dict(spec.full_spec(), 1234), "", 7)
sets = []
for B, seed in ((2, 1), (3, 8), (2, 5)):
    v, w, g = synth.video(B, 29, seed=seed), synth.wav(B, seed=seed), synth.gumbel(B, 29, seed=seed)
    ref, refl = be.infer(v.cuda(), w.cuda(), g.cuda())
    torch.cuda.synchronize()
    sets.append(dict(B=B, pins=[t.pin_memory() for t in (v, w, g)], ref=ref.cpu(), refl=refl.cpu(),
                     out=(torch.empty(B, 80, 300).pin_memory(), torch.empty(B, dtype=torch.int64).pin_memory())))
bad = 0
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
for it in range(12):
    if mode in ("sync", "both"):
        for s in sets:
            s["out"][0].zero_()
            be.infer_host(*s["pins"], *s["out"])
            d = float((s["out"][0] - s["ref"]).abs().max())
            if d != 0.0: bad += 1; print(f"iter {it} sync B={s['B']} maxdiff {d:.3e}", flush=True)
    if mode in ("async", "both"):
        order = [(0, sets[0]), (1, sets[1]), (0, sets[2]), (1, sets[0])]
        pending = []
        for slot, s in order:
            if len(pending) == 2:
                ps, pset = pending.pop(0)
                be.infer_host_wait(ps)
                d = float((pset["out"][0] - pset["ref"]).abs().max())
                if d != 0.0: bad += 1; print(f"iter {it} async slot {ps} B={pset['B']} maxdiff {d:.3e}", flush=True)
            s["out"][0].zero_()
            be.infer_host_submit(slot, *s["pins"], *s["out"])
            pending.append((slot, s))
        for ps, pset in pending:
            be.infer_host_wait(ps)
            d = float((pset["out"][0] - pset["ref"]).abs().max())
            if d != 0.0: bad += 1; print(f"iter {it} async(tail) slot {ps} B={pset['B']} maxdiff {d:.3e}", flush=True)
if mode in ("reuse", "both"):
    for it in range(25):
        a, b = sets[0], sets[1]
        a["out"][0].zero_(); b["out"][0].zero_()
        be.infer_host_submit(0, *a["pins"], *a["out"])
        be.infer_host_submit(1, *b["pins"], *b["out"])
        be.infer_host_submit(0, *a["pins"], *a["out"])          # slot 0 again, nothing waited yet
        be.infer_host_wait(1); be.infer_host_wait(0)
        for name, s in (("a", a), ("b", b)):
            d = float((s["out"][0] - s["ref"]).abs().max())
            if d != 0.0: bad += 1; print(f"reuse iter {it} {name} B={s['B']} maxdiff {d:.3e}", flush=True)
        # device path right after, as the pytest does
        m2, _ = be.infer(sets[1]["pins"][0].cuda(), sets[1]["pins"][1].cuda(), sets[1]["pins"][2].cuda())
        d = float((m2.cpu() - sets[1]["ref"]).abs().max())
        if d != 0.0: bad += 1; print(f"reuse iter {it} device-path B=3 maxdiff {d:.3e}", flush=True)
print("mismatches:", bad)
