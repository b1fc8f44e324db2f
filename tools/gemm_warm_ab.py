"""The dense products of the PPI-shaped epoch, isolated, with the instruction-cache warm-up warp of the TMA GEMM on / off
(PLAGNN_TMA_WARM, read per launch).  Each product is launched with the epilogue the engine gives it (bias + act forward,
gate backward, weight gradient with the bias-gradient column) and event-timed in two ways: back to back (the kernel's code is
hot from the previous launch) and interleaved with an unrelated kernel (an aggregation) as in the real epoch.
    python tools/gemm_warm_ab.py > gpurun_out/gemm_warm_ab.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plagnn_b200 import ops

dev = torch.device("cuda:0")
N = 24041
torch.manual_seed(0)


def mat(r, c):
    x = ops.alloc(r, c, dev)
    x.copy_(torch.randn(r, c, device=dev))
    return x


# (name, builder) — builder returns a zero-argument launcher
def fwd(n, k, pairs, act):
    a = [mat(N, k) for _ in range(pairs)]
    w = [mat(n, k) for _ in range(pairs)]
    b = torch.randn(n, device=dev)
    out = ops.alloc(N, n, dev)
    return lambda: ops.gemm(N, n, [(a[i], 0, w[i], 0, k) for i in range(pairs)], bias=b, act=act, out=out, backend=ops.GEMM_TMA)


def dgrad(n, k, pairs, gated):
    a = [mat(N, k) for _ in range(pairs)]
    w = [mat(k, n) for _ in range(pairs)]
    g = mat(N, n) if gated else None
    out = ops.alloc(N, n, dev)
    return lambda: ops.gemm(N, n, [(a[i], 0, w[i], 1, k) for i in range(pairs)], gate=g,
                            gate_act=ops.ACT_LEAKY if gated else ops.ACT_NONE, out=out, backend=ops.GEMM_TMA)


def wgrad(m, n):
    dz, x = mat(N, m), mat(N, n)
    dw, db = torch.empty(m, n, device=dev), torch.empty(m, device=dev)
    return lambda: ops.gemm_wgrad_bias(dz, x, dw, db)


products = [("fwd 503x503 relu", fwd(503, 503, 1, ops.ACT_RELU)), ("fwd 400x(503+503) leaky", fwd(400, 503, 2, ops.ACT_LEAKY)),
            ("fwd 400x400 relu", fwd(400, 400, 1, ops.ACT_RELU)), ("fwd 300x(400+400) leaky", fwd(300, 400, 2, ops.ACT_LEAKY)),
            ("fwd 200x(300+300) leaky", fwd(200, 300, 2, ops.ACT_LEAKY)), ("fwd 100x200 leaky", fwd(100, 200, 1, ops.ACT_LEAKY)),
            ("dgrad 503x400", dgrad(503, 400, 1, False)), ("dgrad 400x(300+400) gate", dgrad(400, 350, 2, True)),
            ("dgrad 300x(200+300) gate", dgrad(300, 250, 2, True)), ("dgrad 200x100 gate", dgrad(200, 100, 1, True)),
            ("wgrad 400x503", wgrad(400, 503)), ("wgrad 503x503", wgrad(503, 503)), ("wgrad 300x400", wgrad(300, 400)),
            ("wgrad 200x300", wgrad(200, 300)), ("wgrad 100x200", wgrad(100, 200))]

# an unrelated kernel between two GEMM launches, as in the epoch (evicts the GEMM's code from the SMs' instruction caches)
zx = mat(N, 503)
other = lambda: ops.act_backward(zx, zx, ops.ACT_LEAKY)


def timed(fn, reps, spacer):
    for _ in range(3):
        fn()
        if spacer:
            other()
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(reps):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        if spacer:
            other()
        evs.append((s, t))
    torch.cuda.synchronize()
    for s, t in evs:
        tot += s.elapsed_time(t)
    return tot / reps


out = {}
for warm in ("1", "0", "1", "0"):
    os.environ["PLAGNN_TMA_WARM"] = warm
    for name, fn in products:
        rec = out.setdefault(name, {})
        rec.setdefault(f"warm{warm}_back_to_back_ms", []).append(round(timed(fn, 30, False), 5))
        rec.setdefault(f"warm{warm}_interleaved_ms", []).append(round(timed(fn, 30, True), 5))
os.environ.pop("PLAGNN_TMA_WARM", None)
tot = {k: 0.0 for k in ("warm1_back_to_back_ms", "warm0_back_to_back_ms", "warm1_interleaved_ms", "warm0_interleaved_ms")}
for name, rec in out.items():
    for k in tot:
        tot[k] += min(rec[k])
    print(name, {k: min(v) for k, v in rec.items()}, file=sys.stderr, flush=True)
out["sum_of_best"] = {k: round(v, 5) for k, v in tot.items()}
print(json.dumps(out))
