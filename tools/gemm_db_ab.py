"""The directly written dense products of the PPI-shaped epoch on gemm_tma_kernel (256 x 256 tiles) and on gemm_tma_db_kernel
(256 x 128 tiles, accumulators double-buffered in TMEM): PLAGNN_TMA_DB_NOW = 0 / 1, read per launch.  Results against float64
and against each other, event-timed back to back.
    python tools/gemm_db_ab.py > gpurun_out/gemm_db_ab.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plagnn_b200 import ops

dev = torch.device("cuda:0")
N = int(os.environ.get("PLAGNN_AB_ROWS", "24041"))
torch.manual_seed(0)


def mat(r, c):
    x = ops.alloc(r, c, dev)
    x.copy_(torch.randn(r, c, device=dev))
    return x


def act_ref(y, act):
    if act == ops.ACT_RELU:
        return torch.relu(y)
    if act == ops.ACT_LEAKY:
        return torch.where(y > 0, y, y * 0.01)
    return y


def fwd(n, k, pairs, act):
    a = [mat(N, k) for _ in range(pairs)]
    w = [mat(n, k) for _ in range(pairs)]
    b = torch.randn(n, device=dev)
    out = ops.alloc(N, n, dev)
    run = lambda: ops.gemm(N, n, [(a[i], 0, w[i], 0, k) for i in range(pairs)], bias=b, act=act, out=out, backend=ops.GEMM_TMA)
    ref = lambda: act_ref(sum(a[i].double() @ w[i].double().t() for i in range(pairs)) + b.double(), act)
    return run, ref, out


def dgrad(n, k, pairs, gated):
    a = [mat(N, k) for _ in range(pairs)]
    w = [mat(k, n) for _ in range(pairs)]
    g = mat(N, n) if gated else None
    out = ops.alloc(N, n, dev)
    run = lambda: ops.gemm(N, n, [(a[i], 0, w[i], 1, k) for i in range(pairs)], gate=g,
                           gate_act=ops.ACT_LEAKY if gated else ops.ACT_NONE, out=out, backend=ops.GEMM_TMA)

    def ref():
        y = sum(a[i].double() @ w[i].double() for i in range(pairs))
        return y * torch.where(g.double() > 0, 1.0, 0.01) if gated else y
    return run, ref, out


products = [("fwd 503x503 relu", fwd(503, 503, 1, ops.ACT_RELU)), ("fwd 400x(503+503) leaky", fwd(400, 503, 2, ops.ACT_LEAKY)),
            ("fwd 400x400 relu", fwd(400, 400, 1, ops.ACT_RELU)), ("fwd 300x(400+400) leaky", fwd(300, 400, 2, ops.ACT_LEAKY)),
            ("fwd 300x300 relu", fwd(300, 300, 1, ops.ACT_RELU)), ("fwd 200x(300+300) leaky", fwd(200, 300, 2, ops.ACT_LEAKY)),
            ("fwd 100x200 leaky", fwd(100, 200, 1, ops.ACT_LEAKY)), ("fwd 64x72 none", fwd(64, 72, 1, ops.ACT_NONE)),
            ("dgrad 503x400", dgrad(503, 400, 1, False)), ("dgrad 400x300", dgrad(400, 300, 1, False)),
            ("dgrad 300x200", dgrad(300, 200, 1, False)),
            ("dgrad 400x(300+400) gate", dgrad(400, 350, 2, True)), ("dgrad 300x(200+300) gate", dgrad(300, 250, 2, True)),
            ("dgrad 200x100 gate", dgrad(200, 100, 1, True))]


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / reps


out = {}
VARIANTS = ("0", "auto", "two", "db2")
tot = {v: 0.0 for v in VARIANTS}
for name, (run, ref, o) in products:
    rec = {}
    r64 = ref()
    res = {}
    for db in VARIANTS:
        os.environ["PLAGNN_GEMM_PARITY"] = "2" if db in ("two", "db2") else "0"
        os.environ["PLAGNN_TMA_DB2"] = "0" if db == "two" else "1"
        os.environ["PLAGNN_TMA_DB_NOW"] = {"0": "0", "auto": "auto"}.get(db, "1")
        os.environ["PLAGNN_TMA_DB_RING"] = {"r43": "0", "r44": "2"}.get(db, "1")
        o.zero_()
        run()
        torch.cuda.synchronize()
        res[db] = o.clone()
        rec[f"db{db}_err_vs_f64"] = float((res[db].double() - r64).abs().max() / r64.abs().max())
        rec[f"db{db}_ms"] = round(min(timed(run), timed(run)), 5)
        tot[db] += rec[f"db{db}_ms"]
    rec["max_diff_vs_db0"] = max(float((res[v] - res["0"]).abs().max() / res["0"].abs().max()) for v in VARIANTS)
    rec["db2_vs_two"] = float((res["db2"] - res["two"]).abs().max() / res["two"].abs().max())
    out[name] = rec
    print(name, {k: (round(v, 9) if isinstance(v, float) and v < 1e-3 else v) for k, v in rec.items() if k.endswith("_ms") or k in ("max_diff_vs_db0", "db2_vs_two", "dbdb2_err_vs_f64", "dbtwo_err_vs_f64")}, file=sys.stderr, flush=True)
os.environ.pop("PLAGNN_TMA_DB_NOW", None); os.environ.pop("PLAGNN_TMA_DB_RING", None); os.environ.pop("PLAGNN_GEMM_PARITY", None); os.environ.pop("PLAGNN_TMA_DB2", None)
out["sum_ms"] = {k: round(v, 5) for k, v in tot.items()}
print("sum", out["sum_ms"], file=sys.stderr)
print(json.dumps(out))
