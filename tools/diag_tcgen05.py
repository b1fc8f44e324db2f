"""GPU diagnostic for the tcgen05 GEMM backend: small shapes first, error patterns printed.
Usage (on the GPU box): python tools/diag_tcgen05.py > gpurun_out/diag_tcgen05.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plagnn_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(m, n, k, at=0, bt=0, pattern="rand"):
    if pattern == "rand":
        A = torch.randn(m, k)
        B = torch.randn(n, k)
    elif pattern == "index":     # A[i,kk] = 1 only at kk == i % k; B[j,kk] = j + 1000*kk -> C[i,j] = j + 1000*(i%k)
        A = torch.zeros(m, k)
        A[torch.arange(m), torch.arange(m) % k] = 1.0
        B = (torch.arange(n).float().unsqueeze(1) + 1000.0 * torch.arange(k).float().unsqueeze(0))
    a = ops.aligned((A.t().contiguous() if at else A).to(dev))
    b = ops.aligned((B.t().contiguous() if bt else B).to(dev))
    want = (A.double() @ B.double().t())
    got = ops.gemm(m, n, [(a, at, b, bt, k)], backend=ops.GEMM_TCGEN05)
    torch.cuda.synchronize()
    got = got.cpu().double()
    err = ((got - want).abs().max() / want.abs().max()).item()
    print(f"m={m} n={n} k={k} at={at} bt={bt} {pattern}: rel err {err:.3e}", flush=True)
    if err > 1e-5:
        bad = (got - want).abs() > 1e-4 * want.abs().max()
        print("  bad fraction", bad.float().mean().item(), "bad rows", bad.any(1).nonzero().flatten()[:16].tolist(),
              "bad cols", bad.any(0).nonzero().flatten()[:16].tolist())
        print("  got[:4,:8]\n", got[:4, :8], "\n  want[:4,:8]\n", want[:4, :8])
    return err


if __name__ == "__main__":
    errs = []
    errs.append(run(128, 128, 32, pattern="index"))
    errs.append(run(128, 128, 32))
    errs.append(run(128, 128, 8))
    errs.append(run(128, 128, 128))
    errs.append(run(128, 64, 64))
    errs.append(run(100, 48, 40))
    errs.append(run(256, 256, 96))
    errs.append(run(1000, 503, 503))
    errs.append(run(128, 128, 64, at=1, bt=1, pattern="index"))
    errs.append(run(300, 200, 1000, at=1, bt=1))
    errs.append(run(300, 200, 100, at=0, bt=1))
    errs.append(run(128, 128, 32, at=0, bt=1, pattern="index"))
    errs.append(run(128, 128, 32, at=1, bt=0, pattern="index"))
    errs.append(run(260, 136, 70, at=1, bt=0))
    errs.append(run(400, 503, 24041, at=1, bt=1))
    print("max err", max(errs))
