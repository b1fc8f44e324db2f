"""One torchrun launch, many partition configurations on the same generated graph (GPU-minutes: process start-up on 8 GPUs costs
more than a configuration).  Sweeps NCCL CTA caps x row chunks x partition mode x reducer; one JSON line per configuration.

    python -m torch.distributed.run --nproc-per-node 8 ... tools/dist_sweep.py [--steps 5] [--ctas 0,16] [--chunks 2,4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from plagnn_b200 import dist_bench, synth
from plagnn_b200.dist import NcclComm, P2PExchange, block_bounds_by_edges

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--ctas", default="0,16")
ap.add_argument("--chunks", default="2,4")
ap.add_argument("--modes", default="rows,cols")
ap.add_argument("--reducers", default="sum,max")
ap.add_argument("--exchange", default="p2p,nccl", help="feature partition: peer-memory exchange and / or NCCL all-to-all")
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--edges", type=int, default=100_000_000)
ap.add_argument("--feat", type=int, default=256)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, e, f = args.nodes, args.edges, args.feat
if rank == 0:
    sg = synth.scaled_graph(n, e, seed=1234, device=dev)
    h0 = torch.randn(n, f, generator=torch.Generator(device=dev).manual_seed(100), device=dev)
    ne = torch.tensor([sg.src.numel()], device=dev)
else:
    sg, h0, ne = None, torch.empty((n, f), device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
dist.broadcast(ne, 0)
if rank != 0:
    k = int(ne.item())
    sg = synth.ScaledGraph(torch.empty(k, dtype=torch.int64, device=dev), torch.empty(k, dtype=torch.int64, device=dev),
                           torch.empty(k, dtype=torch.float32, device=dev), n)
for t in (sg.src, sg.dst, sg.weight, h0):
    dist.broadcast(t, 0)
refs = {}
for reducer in args.reducers.split(","):
    if rank == 0:
        ref = dist_bench.Variant(sg, n, f, "rows", reducer, 0, 1, dev, None, 1, "rows", h0)
        out, grads = ref.forward_backward()
        refs[reducer] = (out[:n].clone(), [g.clone() for g in grads])
        del ref, out, grads
        torch.cuda.empty_cache()
    dist.barrier()
p2p = None
if "p2p" in args.exchange.split(",") and "cols" in args.modes.split(","):
    per = block_bounds_by_edges(torch.bincount(sg.dst, minlength=n), world)[0]
    try:
        p2p = P2PExchange(per * f * 4, rank, world, dev)
    except Exception as ex:
        if rank == 0:
            print(json.dumps({"p2p_unavailable": repr(ex)[:300]}), flush=True)
for ctas in [int(c) for c in args.ctas.split(",")]:
    comm = NcclComm(rank, world, dev, max_ctas=ctas)
    probe = dist_bench.probe_collectives(comm, n, f, world, dev)
    if rank == 0:
        print(json.dumps({"nccl_max_ctas": ctas, "nccl_alone": probe}), flush=True)
    for reducer in args.reducers.split(","):
        for mode in args.modes.split(","):
            for chunks, xp in ([(int(c), None) for c in args.chunks.split(",")] if mode == "rows" else
                               [(1, p2p if x == "p2p" else None) for x in args.exchange.split(",") if x == "nccl" or p2p is not None]):
                v = dist_bench.Variant(sg, n, f, mode, reducer, rank, world, dev, comm, chunks, "edges", h0, p2p=xp)
                r = dist_bench.measure_variant(v, args.steps, 3, rank, world, dev, refs.get(reducer))
                if rank == 0:
                    r.update({"nccl_max_ctas": ctas, "chunks": chunks, "n_gpus": world})
                    r.pop("kernels_full", None)
                    print(json.dumps(r), flush=True)
                del v
                torch.cuda.empty_cache()
    torch.cuda.synchronize()
    comm.destroy()
if p2p is not None:
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"p2p_wait_gave_up_at_seq": p2p.error()}), flush=True)
    p2p.destroy()
dist.destroy_process_group()
