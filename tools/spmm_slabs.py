"""Does processing the max aggregation in column slabs (smaller per-row footprint -> more source rows in L1) pay?
Calls plagnn_spmm_max_fwd on column sub-ranges of the same matrix (same pitches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plagnn_b200 as P
from plagnn_b200 import ops, synth, _lib
from plagnn_b200.ops import _p, _stream, check, alloc, workspace, REDUCE_MAX

dev = torch.device("cuda:0")
prob = synth.ppi_problem(state="inter")
N = prob.num_nodes
g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=N).add_self_loop().to(dev)
csc = g.csc()
lib = _lib.load()


def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


for F in (503, 400, 300):
    x = ops.aligned(torch.relu(torch.randn(N, F, device=dev)))
    out = alloc(N, F, dev); arg = alloc(N, F, dev, dtype=torch.int32)
    ref_out, ref_arg = ops.spmm_max_fwd(csc, x)
    for slab in (F, 256, 128, 64):
        nb = lib.plagnn_spmm_partial_bytes(csc.counts[2], min(slab, F), REDUCE_MAX)
        part = workspace(nb, dev, "spmm_partial")
        def run():
            c0 = 0
            while c0 < F:
                w = min(slab, F - c0)
                check(lib.plagnn_spmm_max_fwd(_p(csc.indptr), _p(csc.indices), _p(csc.plan), csc.counts_c, N,
                                              x.data_ptr() + 4 * c0, x.stride(0), w, out.data_ptr() + 4 * c0, arg.data_ptr() + 4 * c0,
                                              out.stride(0), _p(part), nb, _stream()), "spmm")
                c0 += w
        ms = timeit(run)
        ok = torch.equal(out[:, :F], ref_out[:, :F]) and torch.equal(arg[:, :F], ref_arg[:, :F])
        print(f"F={F} slab={slab}: {ms:.4f} ms  identical={ok}", flush=True)
