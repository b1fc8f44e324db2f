"""The partitioned workload of BASELINE.json configs[3] (synthetic power-law graph, 1 M nodes / 100 M weighted edges /
256-d features by default; 2-layer weighted-sum GCN and 2-layer SAGEConv-pool stack) on 1/2/4/8 GPUs.

Used two ways: `bench.py --workload scaled` prints it as the line's own metric, and every default `bench.py --gpus N` run
carries the same measurement as the `partitioned` block beside the PPI-shaped `value` (so the driver's 1 -> 8 scaling run
measures the split the north_star asks for, not only independent replicas).

For N > 1 every variant is checked numerically inside the run: rank 0 repeats the first step on the whole graph (world 1) and
the gathered output rows and the all-reduced gradients of the N-rank run must agree with it to 1e-5.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist

CHECK_TOL = 1e-5          # output rows of the first step against the whole-graph run
# Gradients: the row partition reproduces the whole-graph forward pass bit for bit (measured: output error 0.0, gradients
# 1.7e-6 .. 3.5e-6 from the different split of the weight-gradient sums).  The feature partition's narrow kernel adds a row's
# in-edges in another order (output error 4e-7); a handful of pre-activations within that distance of zero then take the other
# leaky_relu branch, and the bias gradients (column sums over 1 M rows that cancel) move by up to 3.5e-5 — the near-tie effect of
# DESIGN.md section 3 again, not an exchange error (a wrong block offset shows up as percents, as the first N = 2 run did).
CHECK_TOL_GRAD = 1e-4


def _rel(a, b):
    d = b.abs().max().item()
    return (a - b).abs().max().item() / d if d > 0 else (a - b).abs().max().item()


def _hbm_peak():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Variant:
    """One (partition mode, reducer) combination on one rank: plan, device structures, model, step function."""

    def __init__(self, sg, n, f, mode, reducer, rank, world, dev, comm, chunks, balance, h0_global, layers=2, p2p=None):
        import plagnn_b200 as P
        from .dist import (CudaBackend, DistGCN, DistSAGEPool, FeaturePartitionPlan, PartitionedGraph, RowPartitionPlan,
                           dist_gcn_forward_backward, dist_pool_forward_backward)
        from . import ops
        self.mode, self.reducer, self.world, self.n, self.f = mode, reducer, world, n, f
        if mode == "cols":
            self.plan = FeaturePartitionPlan(sg.src, sg.dst, n, rank, world, balance=balance)
        else:
            self.plan = RowPartitionPlan(sg.src, sg.dst, n, rank, world, chunks, balance=balance)
        plan = self.plan
        self.pg = PartitionedGraph(plan, sg.weight if reducer == "sum" else None, P.build_csr, dev, transposed=reducer == "sum",
                                   feat=f)
        dims = [f] * (layers + 1)
        self.model = (DistGCN(dims, seed=7) if reducer == "sum" else DistSAGEPool(dims, seed=7)).to(dev)
        self.params = self.model.grad_order()
        self.opt = P.FusedAdam(self.params, lr=1e-3)
        self.h0 = ops.alloc(plan.per, f, dev, zero=True)
        self.h0[:plan.n_local].copy_(h0_global[plan.r0:plan.r1])
        self.backend = CudaBackend(self.pg, comm, p2p if mode == "cols" else None)
        self.exchange = "p2p" if (mode == "cols" and p2p is not None) else ("nccl" if world > 1 else None)
        self.inv = 1.0 / (n * f)
        self._fb = dist_gcn_forward_backward if reducer == "sum" else dist_pool_forward_backward
        self.layers = layers

    def loss_grad(self, out):                            # loss = 0.5 * mean(out^2) over the real rows (padding masked by the library)
        return out * self.inv

    def forward_backward(self):
        return self._fb(self.model, self.pg, self.h0, self.backend, None, self.loss_grad)

    def step(self):
        out, grads = self.forward_backward()
        for p, g in zip(self.params, grads):
            p.grad = g.contiguous()
        self.opt.step()
        return out


def _gather_rows(out_local, plan, world, dev):
    """Output rows of every rank in node order on rank 0 (check plumbing: torch.distributed, not the data path)."""
    if world == 1:
        return out_local[:plan.n_local].clone()
    mine = out_local.contiguous()
    buf = torch.empty((world * mine.shape[0], mine.shape[1]), device=dev, dtype=torch.float32)
    dist.all_gather_into_tensor(buf, mine)
    per = mine.shape[0]
    return torch.cat([buf[r * per: r * per + (plan.bounds[r + 1] - plan.bounds[r])] for r in range(world)])


def measure_variant(v: Variant, steps, warmup, rank, world, dev, reference=None):
    """Times `steps` steps of one variant (max over ranks), the same steps with per-call events (per-collective and per-kernel
    times) and the same steps without the collectives (exposed exchange time = difference).  reference: (out, grads) of the
    whole-graph run of the FIRST step on rank 0, compared before any optimiser step."""
    from . import ops

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / k

    check = None
    barrier()            # ranks build their structures at different speeds; the peer-memory wait gives up after a few seconds
    out, grads = v.forward_backward()
    full = _gather_rows(out, v.plan, world, dev)
    csum = {"out_sum": float(full.double().sum().item()) if rank == 0 else None,
            "grad_abs_sum": float(sum(g.double().abs().sum().item() for g in grads))}
    if reference is not None and rank == 0:
        ref_out, ref_grads = reference
        errs = [_rel(full, ref_out)] + [_rel(g, r) for g, r in zip(grads, ref_grads)]
        check = {"against": "the same first step on the whole graph (world 1) on rank 0", "out_rel_err": errs[0],
                 "grad_rel_err_max": max(errs[1:]), "grad_rel_err": [float(f"{e:.3g}") for e in errs[1:]],
                 "tol_out": CHECK_TOL, "tol_grad": CHECK_TOL_GRAD,
                 "ok": bool(errs[0] <= CHECK_TOL and max(errs[1:]) <= CHECK_TOL_GRAD)}
    del full
    for _ in range(max(warmup, 3)):
        v.step()
    ms = timed(v.step, steps)
    ops.profile_start()
    ms_prof = timed(v.step, steps)
    prof = ops.profile_stop()
    ms_nocomm = None
    if world > 1:
        v.backend.skip_comm = True
        v.step()
        ms_nocomm = timed(v.step, steps)
        v.backend.skip_comm = False
    if rank != 0:
        return None
    per_step = {}
    for k, (c, t) in prof.items():
        per_step[k[0]] = per_step.get(k[0], 0.0) + t / steps
    coll = {k: round(t, 4) for k, t in per_step.items() if k.startswith("nccl_") or k.startswith("p2p_")}
    # aggregation launches by (kernel, rows of the launch): the forward walks few long rows, the transposed backward of the row
    # partition walks every source row with 1 / world of its out-edges
    agg_shapes = {}
    for k, (c, t) in prof.items():
        if k[0].startswith("spmm"):
            key = f"{k[0]}/rows={k[2]}"
            agg_shapes[key] = round(agg_shapes.get(key, 0.0) + t / steps, 4)
    n, f, plan = v.n, v.f, v.plan
    el = plan.num_local_edges
    agg_names = ("spmm_sum", "spmm_max_fwd", "spmm_max_bwd")
    agg_ms = sum(t for k, t in per_step.items() if k in agg_names)
    fc = f // world if v.mode == "cols" else f
    rows_out = plan.n_padded if v.mode == "cols" else plan.per
    if v.reducer == "sum":          # forward + transposed backward per layer: gathered rows + index/weight + output
        alg = v.layers * 2 * (4 * fc * el + 8 * el + 4 * fc * rows_out)
    else:                           # forward: gathered rows + index + output + arg; backward: dz, arg, zero + write dx
        bwd = 16 * fc * plan.n_padded if v.mode == "cols" else 4 * f * (2 * plan.per + 2 * plan.n_padded)
        alg = v.layers * ((4 * fc * el + 4 * el + 8 * fc * rows_out) + bwd)
    hbm, src = _hbm_peak()
    # ncu DRAM bytes per launch of the forward aggregation kernel of this very shape (one GPU, 1 M / 100 M / 256), from the
    # committed capture; the live launch time comes from this run's events
    dram = None
    tpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ncu_traffic_final.json")
    if world == 1 and n == 1_000_000 and f == 256 and plan.num_local_edges == 100_000_000 and os.path.exists(tpath):
        tj = json.load(open(tpath)).get("spmm_kernel<1, 2, 4, 0, 2>/grid534712" if v.reducer == "sum" else "spmm_kernel<0, 2, 4, 0, 2>/grid534712")
        fwd_name = "spmm_sum" if v.reducer == "sum" else "spmm_max_fwd"
        launches = v.layers * (2 if v.reducer == "sum" else 1)
        if tj and per_step.get(fwd_name):
            ms_launch = per_step[fwd_name] / launches
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            dram = {"kernel": "spmm_kernel<%s, 2, 4> (F = 256, 100 M edges)" % ("sum" if v.reducer == "sum" else "max"),
                    "traffic": traffic, "avg_ms": ms_launch, "frac_dram": traffic / (ms_launch * 1e-3) / 1e9 / hbm,
                    "traffic_source": "ncu --set full capture committed under profiles/r2_ncu_final.md (per launch)"}
    res = {"mode": v.mode, "reducer": v.reducer, "exchange": v.exchange, "ms_per_step": ms, "ms_per_step_with_events": ms_prof,
           "ms_per_step_without_collectives": ms_nocomm, "exposed_exchange_ms": (ms - ms_nocomm) if ms_nocomm else 0.0,
           "collective_ms_per_step": coll, "aggregation_ms_per_step": agg_ms, "aggregation_by_launch_shape": agg_shapes,
           "aggregation_algorithmic_gbs_per_gpu": alg / (agg_ms * 1e-3) / 1e9 if agg_ms else None,
           "aggregation_frac_of_hbm_peak": alg / (agg_ms * 1e-3) / 1e9 / hbm if agg_ms else None, "hbm_peak": hbm, "peak_source": src,
           "aggregation_dram": dram,
           "local_edges_rank0": el, "source_slabs": v.pg.n_slabs, "checksum": csum, "check": check,
           "kernels": sorted([{"kernel": k, "ms_per_step": round(t, 4)} for k, t in per_step.items()], key=lambda d: -d["ms_per_step"])[:8]}
    return res


def probe_collectives(comm, n, f, world, dev, reps=5):
    """The exchange steps alone (nothing else on the GPU), at the sizes one layer of the workload moves: what NCCL delivers on
    this box through the library's wrappers.  GB/s = bytes a rank receives from its peers / time (max over ranks)."""
    from . import _lib
    lib = _lib.load()
    per = (n + world - 1) // world
    st = torch.cuda.current_stream().cuda_stream
    local = torch.randn(per, f, device=dev)
    full = torch.empty(world * per, f, device=dev)
    blocks = torch.empty(world * per, f // world, device=dev)
    blocks2 = torch.empty_like(blocks)
    calls = {
        "allgather_rows": (lambda: lib.plagnn_nccl_allgather_rows(local.data_ptr(), full.data_ptr(), per, f, comm.handle, st),
                           4.0 * per * f * (world - 1)),
        "reducescatter_rows": (lambda: lib.plagnn_nccl_reducescatter_rows(full.data_ptr(), local.data_ptr(), per, f, comm.handle, st),
                               4.0 * per * f * (world - 1)),
        "alltoall_blocks": (lambda: lib.plagnn_nccl_alltoall_blocks(blocks.data_ptr(), blocks2.data_ptr(), per * (f // world), world,
                                                                    comm.handle, st), 4.0 * per * (f // world) * (world - 1)),
    }
    out = {}
    for name, (fn, nbytes) in calls.items():
        for _ in range(2):
            _lib.check(fn(), name)
        dist.barrier()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(reps):
            _lib.check(fn(), name)
        e_.record()
        torch.cuda.synchronize()
        t = torch.tensor([s_.elapsed_time(e_) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"ms": round(t.item(), 4), "received_gb_per_s": round(nbytes / (t.item() * 1e-3) / 1e9, 1)}
    return out


def run_partitioned(n, e, f, steps, warmup, rank, world, dev, modes=("rows", "cols"), reducers=("sum", "max"), chunks=None,
                    balance="edges", max_ctas=None):
    """All requested variants on the same generated graph.  Returns (dict for the JSON line on rank 0, None elsewhere)."""
    from . import synth
    from .dist import NcclComm
    t0 = time.perf_counter()
    # Generated on rank 0's GPU and broadcast: device-side generation (randperm / unique over 1e8 keys) is not guaranteed
    # to be bit-identical from one GPU to the next, and a rank working on a slightly different graph was what the first
    # N = 2 run of the in-run check caught (a few output rows off by percents, gradients by 1e-4).
    if rank == 0 or world == 1:
        sg = synth.scaled_graph(n, e, seed=1234, device=dev)
        gen = torch.Generator(device=dev).manual_seed(100)
        h0_global = torch.randn(n, f, generator=gen, device=dev)
        num_e = torch.tensor([sg.src.numel()], device=dev, dtype=torch.int64)
    else:
        sg, h0_global, num_e = None, torch.empty((n, f), device=dev), torch.zeros(1, device=dev, dtype=torch.int64)
    if world > 1:
        dist.broadcast(num_e, 0)
        if rank != 0:
            ne = int(num_e.item())
            sg = synth.ScaledGraph(torch.empty(ne, dtype=torch.int64, device=dev), torch.empty(ne, dtype=torch.int64, device=dev),
                                   torch.empty(ne, dtype=torch.float32, device=dev), n)
        for t in (sg.src, sg.dst, sg.weight, h0_global):
            dist.broadcast(t, 0)
    t_gen = time.perf_counter() - t0
    chunks = chunks or int(os.environ.get("PLAGNN_DIST_CHUNKS", "2" if world > 1 else "1"))
    max_ctas = int(os.environ.get("PLAGNN_NCCL_MAX_CTAS", "0")) if max_ctas is None else max_ctas
    comm = NcclComm(rank, world, dev, max_ctas=max_ctas) if world > 1 else None
    probe = probe_collectives(comm, n, f, world, dev) if world > 1 else None
    p2p = None
    if world > 1 and "cols" in modes and f % (4 * world) == 0 and os.environ.get("PLAGNN_DIST_P2P", "1") != "0":
        from .dist import P2PExchange, block_bounds_by_edges
        per = block_bounds_by_edges(torch.bincount(sg.dst, minlength=n), world)[0] if balance == "edges" else (n + world - 1) // world
        try:
            p2p = P2PExchange(per * f * 4, rank, world, dev)
        except Exception as ex:                       # e.g. CUDA IPC not permitted in this container: NCCL all-to-all instead
            if rank == 0:
                import sys
                print(f"[plagnn] peer-memory exchange unavailable ({ex!r}); using the NCCL all-to-all", file=sys.stderr, flush=True)
            p2p = None
    results = []
    single_ms = {}
    for reducer in reducers:
        reference = None
        if world > 1:
            if rank == 0:                                # the whole-graph run of the first step, before any optimiser step
                ref = Variant(sg, n, f, "rows", reducer, 0, 1, dev, None, 1, "rows", h0_global)
                out, grads = ref.forward_backward()
                reference = (out[:n].clone(), [g.clone() for g in grads])
                del out, grads
                # ... and its step time on this GPU alone: the denominator of the strong-scaling efficiency of this run
                for _ in range(3):
                    ref.step()
                torch.cuda.synchronize()
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                for _ in range(steps):
                    ref.step()
                e_.record()
                torch.cuda.synchronize()
                single_ms[reducer] = s_.elapsed_time(e_) / steps
                del ref
                torch.cuda.empty_cache()
            dist.barrier()
        for mode in modes:
            if mode == "cols" and f % (4 * world):
                continue
            v = Variant(sg, n, f, mode, reducer, rank, world, dev, comm, chunks, balance, h0_global, p2p=p2p)
            r = measure_variant(v, steps, warmup, rank, world, dev, reference)
            if r is not None:
                r["chunks"] = v.plan.chunks
                results.append(r)
            del v
            torch.cuda.empty_cache()
        del reference
    p2p_err = None
    if p2p is not None:
        torch.cuda.synchronize()
        p2p_err = p2p.error()
        p2p.destroy()
    if comm is not None:
        torch.cuda.synchronize()
        comm.destroy()
    if rank != 0:
        return None
    layers = 2
    for r in results:
        r["edges_per_s"] = e * layers * 2 / (r["ms_per_step"] * 1e-3)      # forward + backward aggregation, all ranks
        if r["reducer"] in single_ms:
            r["single_gpu_ms_per_step_same_run"] = single_ms[r["reducer"]]
            r["speedup_vs_single_gpu"] = single_ms[r["reducer"]] / r["ms_per_step"]
            r["strong_scaling_efficiency"] = single_ms[r["reducer"]] / r["ms_per_step"] / world
    best = {}
    for r in results:
        k = r["reducer"]
        if k not in best or r["ms_per_step"] < best[k]["ms_per_step"]:
            best[k] = r
    return {"workload": f"BASELINE configs[3]: power-law graph N={n}, E={e} directed weighted edges, F={f}; 2-layer weighted-sum GCN "
                        "(u_mul_e + sum, right-normalised, bias + leaky_relu fused) and 2-layer SAGEConv-pool stack (max reducer); "
                        "step = forward + backward + weight-gradient all-reduce + Adam", "scaling": "strong", "n_gpus": world,
            "balance": balance, "nccl_max_ctas": max_ctas, "generate_seconds": round(t_gen, 2), "nccl_alone": probe,
            "variants": results,
            "winner": {k: {"mode": v["mode"], "ms_per_step": v["ms_per_step"], "edges_per_s": v["edges_per_s"],
                           "strong_scaling_efficiency": v.get("strong_scaling_efficiency")} for k, v in best.items()},
            "p2p_wait_gave_up_at_seq": p2p_err,
            "checks_ok": all(r["check"]["ok"] for r in results if r["check"]) if world > 1 else None}


def run(args):
    """`bench.py --workload scaled`: the partitioned workload as the line's own metric (weighted-sum reducer, best partition)."""
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
    modes = tuple(os.environ.get("PLAGNN_DIST_MODES", "rows,cols").split(","))
    reducers = tuple(os.environ.get("PLAGNN_DIST_REDUCERS", "sum,max").split(","))
    block = run_partitioned(n, e, f, args.steps, args.warmup, rank, world, dev, modes=modes, reducers=reducers)
    if rank == 0:
        first = "sum" if "sum" in block["winner"] else next(iter(block["winner"]))
        w = block["winner"][first]
        wr = next(r for r in block["variants"] if r["reducer"] == first and r["mode"] == w["mode"])
        print(json.dumps({
            "metric": "GCN fwd+bwd epochs/s", "value": 1e3 / w["ms_per_step"], "unit": "epochs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": w["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": block["workload"], "parallelism": f"{w['mode']} partition x{world} ({first} reducer)",
                       "l2": "inputs (>= 1 GB of features per aggregation at N = 1) larger than L2"},
            "spmm_edges_per_s": w["edges_per_s"],
            "roofline": {"kernel": f"aggregation ({first}) F={f}, rank 0 share", "bound": "hbm",
                         "achieved": wr["aggregation_algorithmic_gbs_per_gpu"], "peak": wr["hbm_peak"], "unit": "GB/s",
                         "frac": wr["aggregation_frac_of_hbm_peak"], "traffic": None},
            "gpu_launches": None, "partitioned": block}))
    if world > 1:
        dist.destroy_process_group()
