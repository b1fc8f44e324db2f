"""[diagnostics build] Cycle sums per role of gemm_tma_db_kernel (PLAGNN_TMA_TRACE=1): python tools/gemm_trace_db.py m n k bt [pairs]"""
import ctypes, os, sys
os.environ["PLAGNN_TMA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plagnn_b200 import ops, _lib
dev = torch.device("cuda:0")
m, n, k, bt = (int(v) for v in sys.argv[1:5])
pr = int(sys.argv[5]) if len(sys.argv) > 5 else 1
a = ops.aligned(torch.randn((m, k), device=dev))
b = ops.aligned(torch.randn((k, n) if bt else (n, k), device=dev))
out = ops.alloc(m, n, dev)
for _ in range(3):
    ops.gemm(m, n, [(a, 0, b, bt, k)] * pr, out=out, backend=ops.GEMM_TMA)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (64 * 32))()
lib.plagnn_tma_trace.argtypes = [ctypes.c_void_p]
assert lib.plagnn_tma_trace(buf) == 0
print(f"m={m} n={n} k={k} bt={bt} pairs={pr} ring={os.environ.get('PLAGNN_TMA_DB_RING', '1')}")
print("cta rank tiles kblocks | total | first MMA at | issuer: wait lo_full, wait acc_empty, last issue at | producer wait raw_empty | "
      "split warp 2: wait raw_full, wait lo_empty, work | read-out warp: wait acc_full, read-out")
for c in range(0, 6):
    r = buf[32 * c: 32 * c + 32]
    if r[7] == 0:
        continue
    t0 = r[0]
    print(f"{c:3d} {r[12]:4d} {r[1]:5d} {r[11]:7d} | {r[7]-t0:8d} | {(r[2]-t0) if r[2] else 0:8d} | {r[9]:8d} {r[3]:8d} {(r[4]-t0) if r[4] else 0:8d} | {r[8]:8d} | "
          f"{r[10]:8d} {r[13]:8d} {r[14]:8d} | {r[15]:8d} {r[16]:8d}")
