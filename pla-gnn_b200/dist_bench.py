"""`bench.py --workload scaled`: row-partitioned weighted-sum GCN on the synthetic power-law graph
(BASELINE.json configs[3]: 1 M nodes / 100 M weighted edges / 256-d features by default), 1/2/4/8 GPUs."""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def run(args):
    import plagnn_b200 as P
    from plagnn_b200 import ops, synth
    from plagnn_b200.dist import CudaBackend, DistGCN, PartitionedGraph, RowPartitionPlan, dist_gcn_forward_backward

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.nodes or 1_000_000
    e = args.edges or 100_000_000
    f = args.feat
    t0 = time.perf_counter()
    sg = synth.scaled_graph(n, e, seed=1234, device=dev)                  # identical on every rank (seeded)
    chunks = int(os.environ.get("PLAGNN_DIST_CHUNKS", "4" if world > 1 else "1"))
    plan = RowPartitionPlan(sg.src, sg.dst, n, rank, world, chunks)
    pg = PartitionedGraph(plan, sg.weight, P.build_csr, dev)
    del sg
    torch.cuda.empty_cache()
    t_build = time.perf_counter() - t0
    model = DistGCN([f, f, f], seed=7).to(dev)
    params = [p for pair in zip(model.weights, model.biases) for p in pair]
    opt = P.FusedAdam(params, lr=1e-3)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    h0 = ops.alloc(plan.per, f, dev, zero=True)
    h0[:plan.n_local].copy_(torch.randn(plan.n_local, f, generator=gen, device=dev))
    backend = CudaBackend(pg)
    row_mask = (torch.arange(plan.per, device=dev) < plan.n_local).float().unsqueeze(1)
    inv = 1.0 / (n * f)

    def loss_grad(out):                                  # loss = 0.5 * mean(out^2) over the real rows
        return out * (row_mask * inv)

    def step():
        out, grads = dist_gcn_forward_backward(model, pg, h0, backend, None, loss_grad)
        for p, g in zip(params, grads):
            p.grad = g.contiguous()
        opt.step()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ops.profile_start()
    s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step()
    e_.record()
    barrier()
    prof = ops.profile_stop()
    ms = s.elapsed_time(e_)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        layers = len(model.weights)
        edges_per_step = e * layers * 2                  # forward + transposed backward aggregation, all ranks
        spmm = {k: v for k, v in prof.items() if k[0] == "spmm_sum"}
        spmm_ms = sum(t for (_, t) in spmm.values()) / args.steps
        el = plan.num_local_edges
        alg_local = layers * 2 * (4 * f * el + 4 * el * 2 + 4 * f * plan.per)      # gathered rows + idx/w + out
        hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                          "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
            if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
        print(json.dumps({
            "metric": "GCN fwd+bwd epochs/s", "value": args.steps / (ms * 1e-3), "unit": "epochs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"row-partitioned weighted-sum GCN [{f},{f},{f}] on power-law graph N={n}, E={e} "
                                   f"(BASELINE configs[3]), right-normalised, Adam", "parallelism": f"1-D row partition x{world}, {chunks} row chunks per rank pipelined: "
                       "all-gather(fwd) / reduce-scatter(bwd) / all-reduce(weight grads) over NCCL",
                       "l2": "inputs (>= 1 GB gathered features) larger than L2", "build_seconds": round(t_build, 2),
                       "local_edges_rank0": el},
            "spmm_edges_per_s": edges_per_step / (ms * 1e-3),
            "roofline": {"kernel": f"spmm_sum F={f} (weighted, rank 0 share)", "bound": "hbm",
                         "achieved": alg_local / (spmm_ms * 1e-3) / 1e9 if spmm_ms else None, "peak": hbm, "unit": "GB/s",
                         "frac": (alg_local / (spmm_ms * 1e-3) / 1e9 / hbm) if spmm_ms else None, "traffic": None,
                         "spmm_ms_per_step": spmm_ms},
            "gpu_launches": None,
            "kernels": sorted([{"kernel": "/".join(map(str, k)), "ms_per_step": v[1] / args.steps} for k, v in prof.items()],
                              key=lambda d: -d["ms_per_step"])[:8]}))
    if world > 1:
        dist.destroy_process_group()
