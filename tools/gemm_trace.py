"""[needs the diagnostics build: python pla-gnn_b200/csrc/build.py --diag; run with PLAGNN_LIB_PATH=pla-gnn_b200/libplagnn_diag.so] Per-CTA timeline of one TMA GEMM launch (PLAGNN_TMA_TRACE=1): python tools/gemm_trace.py m n k at bt"""
import ctypes, os, sys
os.environ["PLAGNN_TMA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plagnn_b200 import ops, _lib
dev = torch.device("cuda:0")
m, n, k, at, bt = (int(v) for v in sys.argv[1:6])
a = ops.aligned(torch.randn((k, m) if at else (m, k), device=dev))
b = ops.aligned(torch.randn((k, n) if bt else (n, k), device=dev))
out = ops.alloc(m, n, dev)
for _ in range(3):
    ops.gemm(m, n, [(a, at, b, bt, k)], out=out, backend=ops.GEMM_TMA)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (64 * 32))()
lib.plagnn_tma_trace.argtypes = [ctypes.c_void_p]
assert lib.plagnn_tma_trace(buf) == 0
print(f"m={m} n={n} k={k} at={at} bt={bt} debug={os.environ.get('PLAGNN_TMA_DEBUG','0')}")
print("cta rank nkb | setup  firstfull  mainloop(issue)  acc_wait  epilogue  storewait  teardown | total | waits: producer(empty) mma(full) conv(raw)   [cycles]")
for c in range(0, 12):
    r = buf[32 * c: 32 * c + 32]
    if r[7] == 0:
        continue
    t0 = r[0]
    print(f"{c:3d} {r[12]:4d} {r[11]:4d} | {r[1]-t0:6d} {((r[2]-r[1]) if r[2] else 0):8d} {((r[3]-r[2]) if r[3] else 0):10d} {((r[4]-r[3]) if r[3] else r[4]-r[1]):10d} "
          f"{r[5]-r[4]:8d} {r[6]-r[5]:8d} {r[7]-r[6]:8d} | {r[7]-t0:8d} | {r[8]:8d} {r[9]:8d} {r[10]:8d} | epi: tmem {r[13]} waitread {r[14]} wait+sts+issue {r[15]}")
    print(f"      epilogue of warp 2: tmem load until first use {r[25]}, wait_read+sts {r[26]}, proxy fence {r[27]}, store issue {r[28]}")
    if r[29] or r[30]:
        print(f"      second tile of this pair: tile 0 committed -> first k-block of tile 1 ready +{r[30]}, main loop {r[31]}, read-out {r[29]}")
    if r[12] == 0:
        print(f"      kb0: tma issue +{r[16]-r[1]}, raw seen +{r[17]-r[16]}, split done +{r[18]-r[17]}, issuer sees lo_full +{r[19]-r[18]}"
              f" | kb8: tma issue @{r[20]-r[1]}, raw seen +{r[21]-r[20]}, lo_empty passed +{r[24]-r[21]}, split done +{r[22]-r[24]}, issuer +{r[23]-r[22]}")
