"""Roofline of the train-step tail kernels at the reference's parameter count (38.44 M fp32)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lip2speech_b200 import _lib, build
build.build()
be = _lib.backend(0)
n = (38_436_836 + 3) // 4 * 4
p, g, m, v, vmax = (torch.randn(n, device="cuda") * 0.01 for _ in range(5))
v.abs_(); vmax.abs_()
sq = torch.zeros(1, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peak = 6542.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

t1 = timeit(lambda: be.allreduce_grads(g, 1.0, sq))
t2 = timeit(lambda: be.clip_adamw_step(p, g, m, v, vmax, sq, 1.0, 1e-4, 0.9, 0.999, 1e-8, 1e-6, 7))
mel, post = torch.randn(64, 80, 77, device="cuda"), torch.randn(64, 80, 77, device="cuda")
tgt, gate, gt = torch.randn(64, 80, 77, device="cuda"), torch.randn(64, 77, device="cuda"), torch.zeros(64, 77, device="cuda")
dis = torch.softmax(torch.randn(256, 501, device="cuda"), -1)
t3 = timeit(lambda: be.loss_fwd_bwd(mel, post, gate, dis, tgt, gt))
print(json.dumps({"n_params": n, "hbm_peak_gbs": peak,
                  "grad_scale_sqnorm": {"ms": t1, "bytes": 8 * n, "gbs": 8 * n / t1 / 1e6, "frac": 8 * n / t1 / 1e6 / peak},
                  "clip_adamw": {"ms": t2, "bytes": 40 * n, "gbs": 40 * n / t2 / 1e6, "frac": 40 * n / t2 / 1e6 / peak},
                  "loss_fwd_bwd_B64_M77": {"ms": t3}}))
