#!/usr/bin/env python
"""Benchmark of the PLA-GNN message-passing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ppi|scaled]

A "step" is one epoch of the reference's training loop body (code/train.py:197-205): zero_grad, full-graph
forward, class-weighted loss on the training rows, backward, Adam step — on the PPI-shaped synthetic graph
(BASELINE.json configs[1]: TSA perturbation-state shape, N = 24 041, E = 1.4 M directed + N self-loops,
503-d features, GNN32 503-400-300-200-100-12).  Prints ONE JSON line (rank 0).

N > 1: the PPI graph fits one GPU, so ranks are independent replicas (one graph / model per GPU, no
data-path collective, "scaling": "weak").  `--workload scaled` runs the row-partitioned synthetic
power-law graph (BASELINE.json configs[3]) with the NCCL all-gather / reduce-scatter exchange.

`--impl reference` times the CPU oracle (DGL-equivalent restatement; DGL itself is not installable in this
image) on the host cores, on the same workload, metric and unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "GCN fwd+bwd epochs/s"
UNIT = "epochs/s"
HIDDEN = (400, 300, 200, 100, 12)       # code/train.py:179
LR = 5e-5                               # code/main_normal.py:26


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ppi", choices=["ppi", "scaled"])
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--edges", type=int, default=None)
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--concurrent-models", type=int, default=4)
    ap.add_argument("--no-partitioned", action="store_true", help="skip the configs[3] partitioned block")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the configs[1] end-to-end pipeline sample")
    ap.add_argument("--pipeline-models", type=int, default=2)
    ap.add_argument("--pipeline-epochs", type=int, default=25)
    ap.add_argument("--partitioned-steps", type=int, default=10)
    ap.add_argument("--partitioned-nodes", type=int, default=1_000_000)
    ap.add_argument("--partitioned-edges", type=int, default=100_000_000)
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.  NVML in-process (nvidia_ml_py) when it loads: a
    query is one ioctl, so a 50 ms poll does not disturb the stream (A/B on B200: 2.25-2.27 ms per epoch with and without it); starting an `nvidia-smi -lms` process inside a 100 ms
    timed region does (its start-up enumerates every GPU of the box and was seen to stall launches).  nvidia-smi is the
    fallback when NVML cannot be loaded.  The poll is started by prepare() before the warm-up; only samples taken between
    start() and stop() are reported."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.nv = None
        self.handle = None
        self.rows = []           # (time, sm_mhz, power_w, reasons bitmask)
        self.t0 = self.t1 = None
        self.thread = None
        self.stop_flag = False
        self.p = None
        self.f = None

    def prepare(self):
        self.period = float(os.environ.get("PLAGNN_CLOCK_POLL_MS", "50")) / 1000.0
        if os.environ.get("PLAGNN_CLOCK_SOURCE") == "smi":
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            handle = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv, self.handle = pynvml, handle
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        self.t0 = time.perf_counter()
        if self.nv is not None:
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        self.t1 = time.perf_counter()
        if self.nv is not None:
            time.sleep(0.03)
            self.stop_flag = True
            self.thread.join(timeout=2)
            rows = [r for r in self.rows if self.t0 <= r[0] <= self.t1] or self.rows[-3:]
            if not rows:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            nv = self.nv
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = [n for n, b in bits.items() if any(r[3] & b for r in rows)]
            sm = sorted(r[1] for r in rows)
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "power_w_max": max(r[2] for r in rows),
                    "samples": len(rows), "reasons": reasons, "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any("Active" == r[5 + j].strip() for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons, "source": "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
def make_problem(rank: int, nodes, edges):
    from plagnn_b200 import synth
    n = nodes or synth.PPI_NODES
    e = edges or synth.PPI_EDGES
    # rank 0: TSA perturbation-state graph; other ranks: independently rewired conditions (configs[2])
    prob = synth.ppi_problem(n, e, "inter", seed=70 + 13 * rank)
    rng = np.random.default_rng(12 + rank)
    lab = prob.labelled.copy()
    rng.shuffle(lab)
    train_index = np.sort(lab[: len(lab) * 9 // 10])          # one fold of KFold(10): 90 % of the labelled rows
    return prob, train_index


def algorithmic_spmm_bytes(n, e_prime, f, with_arg=True):
    """4·F·E' gathered rows + 4·F·N output + 4·E' indices + 4·(N+1) indptr [+ 4·F·N arg]."""
    return 4 * f * e_prime + 4 * f * n + 4 * e_prime + 4 * (n + 1) + (4 * f * n if with_arg else 0)


def run_ours(args):
    import plagnn_b200 as P
    from plagnn_b200 import ops
    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    prob, train_index = make_problem(rank, args.nodes, args.edges)
    n = prob.num_nodes
    ids = list(range(n))

    # ---- host (pinned) copies of what one epoch consumes: features, labels, training rows ----------
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(dev)
    csc = g.csc()
    features, labels = g.ndata["feat"], g.ndata["loc"]
    # pinned host images in the device layout (rows padded to 128 bytes) so each H2D is one contiguous copy
    feat_dev, loc_dev = features._base, labels._base
    feat_h = torch.zeros(feat_dev.shape, dtype=torch.float32).pin_memory()
    feat_h[:, :features.shape[1]] = torch.from_numpy(prob.features)
    loc_h = torch.zeros(loc_dev.shape, dtype=torch.float32).pin_memory()
    loc_h[:, :labels.shape[1]] = torch.from_numpy(prob.loc.astype(np.float32))
    idx_h = torch.from_numpy(train_index).pin_memory()
    idx_d = idx_h.to(dev)
    i_weight = P.weight_cal(prob.loc)
    torch.manual_seed(70)
    model = P.GNN32(features.shape[1], *HIDDEN).to(dev)
    opt = P.FusedAdam(model.parameters(), lr=LR)
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()
    logits_h = torch.empty((n, HIDDEN[-1]), dtype=torch.float32).pin_memory()

    def epoch():
        opt.zero_grad()
        logits = model(g, features)
        loss = P.multi_loss_indexed(logits, labels, idx_d, i_weight)
        loss.backward()
        opt.step()
        return logits, loss

    # End-to-end step: the inputs of EVERY step come from pinned host memory and the loss + N x 12 output go back to
    # the host.  The copies run on a second stream into double-buffered device tensors, so the upload of step i+1
    # overlaps the compute of step i (a prefetching input pipeline); all of it is inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    feat_bufs = [feat_dev, torch.zeros_like(feat_dev)]
    loc_bufs = [loc_dev, torch.zeros_like(loc_dev)]
    idx_bufs = [idx_d, torch.zeros_like(idx_d)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]         # upload of buffer b finished
    consumed = [torch.cuda.Event(), torch.cuda.Event()]      # compute that read buffer b finished
    state = {"step": 0, "primed": False}

    def upload(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])              # do not overwrite inputs still being read
            feat_bufs[b].copy_(feat_h, non_blocking=True)
            loc_bufs[b].copy_(loc_h, non_blocking=True)
            idx_bufs[b].copy_(idx_h, non_blocking=True)
            ready[b].record(copy_stream)

    def epoch_e2e():
        b = state["step"] & 1
        if not state["primed"]:
            consumed[0].record(); consumed[1].record()
            upload(b)
            state["primed"] = True
        upload(b ^ 1)                                        # prefetch the next step's inputs
        torch.cuda.current_stream().wait_event(ready[b])
        f_view = feat_bufs[b][:, :features.shape[1]]
        l_view = loc_bufs[b][:, :labels.shape[1]]
        opt.zero_grad()
        logits = model(g, f_view)
        loss = P.multi_loss_indexed(logits, l_view, idx_bufs[b], i_weight)
        loss.backward()
        opt.step()
        consumed[b].record()
        out_dev = logits.detach().contiguous()
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            loss_h.copy_(loss.detach().reshape(1), non_blocking=True)   # D2H: loss and the N x 12 output the loop reads
            logits_h.copy_(out_dev, non_blocking=True)
            out_dev.record_stream(copy_stream)
        state["step"] += 1

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.prepare()
    for _ in range(3):
        epoch()
    # ---- several models of the fold/seed sweep at once on this GPU, one CUDA stream each (reported beside `value`) ----
    # The reference trains 100 independent models per condition one after the other (code/train.py:162-180); their
    # kernels can share the GPU: the aggregation of one model fills the SMs a GEMM tail wave of another leaves idle.
    conc = None
    if args.concurrent_models > 1:
        extra = []
        for i in range(1, args.concurrent_models):
            torch.manual_seed(70 + i)
            m_i = P.GNN32(features.shape[1], *HIDDEN).to(dev)
            extra.append((m_i, P.FusedAdam(m_i.parameters(), lr=LR), torch.cuda.Stream(device=dev)))

        def round_of_epochs():
            epoch()
            for m_i, o_i, s_i in extra:
                with torch.cuda.stream(s_i):
                    o_i.zero_grad()
                    out_i = m_i(g, features)
                    l_i = P.multi_loss_indexed(out_i, labels, idx_d, i_weight)
                    l_i.backward()
                    o_i.step()

        def timed_conc(steps):
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                round_of_epochs()
            for _, _, s_i in extra:
                torch.cuda.current_stream().wait_stream(s_i)
            e.record()
            barrier()
            return s.elapsed_time(e)

        for _ in range(3):
            round_of_epochs()
        ms_conc = timed_conc(args.steps)
        conc = {"models": args.concurrent_models, "value": world * args.concurrent_models * args.steps / (ms_conc * 1e-3),
                "unit": UNIT, "ms_per_round": ms_conc / args.steps,
                "note": "independent models of the sweep, one CUDA stream each, same graph; `value` above is one model alone"}
        del extra

    # ---- per-kernel breakdown: K steps with events around EVERY entry point (~150 event records per epoch, which costs
    # host time); not part of `value`
    for _ in range(3):
        epoch()
    ops.profile_start()
    ms_profiled = timed(epoch, args.steps)
    prof = ops.profile_stop()
    # ---- the same epoch as one CUDA-graph replay (plagnn_b200.TrainStep: no autograd, no host work per launch) -----------
    graph_leg = None
    try:
        torch.manual_seed(70)
        m_g = P.GNN32(features.shape[1], *HIDDEN).to(dev)
        ts = P.TrainStep(m_g, g, features, labels, idx_d, i_weight, lr=LR)
        for _ in range(3):
            ts.step()
        ms_graph = timed(ts.step, args.steps)
        graph_leg = {"value": world * args.steps / (ms_graph * 1e-3), "unit": UNIT, "ms_per_step": ms_graph / args.steps,
                     "launches_per_step": ts.launches_per_epoch, "loss_after": float(ts.loss.item()),
                     "note": "forward + indexed loss + backward + Adam (step count on the device) as four C-ABI calls captured "
                             "once and replayed; every node is a library kernel, PDL edges kept; `value` above stays the "
                             "autograd-driven loop the unchanged train.py runs"}
        del ts, m_g
    except Exception as ex:                    # a capture problem must not take the headline measurement down
        graph_leg = {"value": None, "error": repr(ex)[:300]}
    # ---- BASELINE configs[1] end to end: control + perturbation state trained, merged, scored, ranked (bounded sample) ----
    config2 = None
    if rank == 0 and not args.no_pipeline:
        try:
            from plagnn_b200 import pipeline, synth
            prob_n = synth.ppi_problem(n, len(prob.ppi_row), "normal", seed=70)
            g_n = P.create_graph(prob_n.scipy_ppi(), prob_n.ecc, prob_n.gcn, prob_n.scipy_loc(), prob_n.expr, ids).to(dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rec, _, _ = pipeline.alteration_pipeline(g_n, g, prob.labelled, prob.loc, lr=LR, fold_num=10, epoch_num=args.pipeline_epochs,
                                                     fold_seeds=(12,), model_seed=70, max_models=args.pipeline_models)
            top_rows = rec["row"][:5].tolist()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            config2 = {"seconds": dt, "models_per_state": args.pipeline_models, "epochs_per_model": args.pipeline_epochs,
                       "epochs_total": 2 * args.pipeline_models * args.pipeline_epochs, "ranked_entries": int(rec["row"].numel()),
                       "top_rows": top_rows, "full_sweep_extrapolated_hours": dt / (2 * args.pipeline_models * args.pipeline_epochs)
                                                                                * (2 * 100 * 200) / 3600.0,
                       "note": "wall clock incl. graph capture per model and KFold on the host: train control and TSA state "
                               "(TrainStep), mat_merge of each, (inter - normal) / normal, rank; the reference's full sweep is 100 "
                               "models x 200 epochs per state (code/train.py:162-205, code/main.py:32-48,80-84)"}
            del g_n, rec
        except Exception as ex:
            config2 = {"seconds": None, "error": repr(ex)[:300]}
    # ---- BASELINE configs[3]: the partitioned synthetic graph on the same N GPUs (reported beside `value`) ---------------
    partitioned = None
    if not args.no_partitioned:
        from plagnn_b200 import dist_bench
        torch.cuda.empty_cache()
        partitioned = dist_bench.run_partitioned(args.partitioned_nodes, args.partitioned_edges, args.feat, args.partitioned_steps,
                                                 3, rank, world, dev, modes=("rows",) if world == 1 else ("rows", "cols"))
        torch.cuda.empty_cache()
    # ---- end-to-end timed region (host buffers in, loss + logits out) ------------------------------
    for _ in range(3):
        epoch_e2e()

    def timed_e2e(steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            epoch_e2e()
        torch.cuda.current_stream().wait_stream(copy_stream)   # the last step's results must have reached the host
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = tt.item()
        return ms

    # ---- the same epoch on the single-chain GEMM kernels (PLAGNN_GEMM_PARITY=0, read per launch by the library): the fastest
    # mode, whose gradients sit at 1.5e-5 instead of <= 1e-5 of the fp32 oracle at this size.  Reported beside
    # `value`; `value` and `e2e` are the default (parity) mode.
    prev_parity = os.environ.get("PLAGNN_GEMM_PARITY")
    os.environ["PLAGNN_GEMM_PARITY"] = "0"
    try:
        for _ in range(3):
            epoch()
        ms_fast = timed(epoch, args.steps)
    finally:
        if prev_parity is None:
            os.environ.pop("PLAGNN_GEMM_PARITY", None)
        else:
            os.environ["PLAGNN_GEMM_PARITY"] = prev_parity
    fast_gemm = {"value": world * args.steps / (ms_fast * 1e-3), "unit": UNIT, "ms_per_step": ms_fast / args.steps,
                 "note": "PLAGNN_GEMM_PARITY=0: one accumulation chain per GEMM tile (256 x 256 or double-buffered 256 x 128 tiles by the cost model), "
                         "weight-gradient chains of 40 k-blocks; full-size gradients within 1.5e-5 instead of 1e-5"}

    ms_e2e = timed_e2e(args.steps)

    # ---- device-resident timed region (value): the LAST leg of the process, W warm-up steps right before it --------------
    # (the legs above are timed regions of their own, each with its own warm-up; the two headline legs — end to end and
    # this one — come last, when the process has been launching kernels for seconds and the clocks are up: a first run of
    # this order with the end-to-end leg first measured it at 5.9 ms per step instead of 2.3)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        epoch()
    barrier()                  # communicator set-up (N > 1) happens here, not inside the sampled window
    if sampler:
        sampler.start()
    launches0 = ops.launch_count()
    # CUDA events on the launching stream, recorded inside the library around the aggregation entry points only (12 event
    # records per epoch); the roofline figure comes from these, i.e. from inside the timed region
    ops.profile_start(aggregation_only=True)
    ms_total = timed(epoch, args.steps)
    prof_agg = ops.profile_stop()
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    hbm_peak, tf_peak, peak_src = measured_peaks()
    e_prime = csc.num_edges
    f_in = features.shape[1]
    kernels = []
    for key, (cnt, tot) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        kernels.append({"kernel": "/".join(str(k) for k in key), "calls_per_step": cnt / args.steps,
                        "avg_ms": round(tot / cnt, 5), "share_of_step": round(tot / ms_profiled, 4)})
    cnt, tot = prof_agg[("spmm_max_fwd", f_in, n, 0)]
    spmm_ms = tot / cnt
    alg = algorithmic_spmm_bytes(n, e_prime, f_in)
    achieved = alg / (spmm_ms * 1e-3) / 1e9
    gemm_ms = sum(t for k, (c, t) in prof.items() if k[0] == "gemm") / args.steps
    spmm_all_ms = sum(t for k, (c, t) in prof_agg.items() if k[0].startswith("spmm")) / args.steps
    flops_epoch = dense_flops(n, f_in)
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
    l2_cap = 6300.0 * sm_mhz * 1e6 / 1e9 if sm_mhz else None
    traffic = lts = None
    tpath = os.path.join(ROOT, "profiles", "r2_spmm_traffic_ppi.json")
    if n == 24041 and os.path.exists(tpath):          # dram__bytes_read + write of this kernel from the committed ncu capture
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        lts = tj.get("lts_sectors_read_by_sm")
    rg = roofline_gemm(prof, args.steps, flops_epoch, gemm_ms, tf_peak, peak_src)
    out = {
        "metric": METRIC, "value": world * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PPI-shaped TSA-state graph (BASELINE configs[1]): N={n}, E={e_prime - n} directed + {n} "
                               f"self-loops, F={f_in}, GNN32 {f_in}-400-300-200-100-12, full-graph epoch "
                               "(zero_grad, fwd, indexed weighted-BCE, bwd, Adam)",
                   "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (one graph/model per GPU)",
                   "l2": "no explicit flush: one step touches >1 GB of distinct activations/gradients (> 126 MB L2); the "
                         "aggregation input is produced by the preceding GEMM, as in the real loop",
                   "legs": "in process order: concurrent models (3 warm-up + K), per-kernel breakdown (3 + K), graph replay (3 + K), partitioned "
                           "configs[3] block, single-chain GEMM mode (3 + K, `gemm_mode.fast`), end-to-end (3 + K), then W warm-up + K timed steps = "
                           "`value` (default GEMM mode: two accumulation chains per tile)",
                   "train_rows": int(len(train_index)), "lr": LR},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(feat_h.numel() * 4 + loc_h.numel() * 4 + idx_h.numel() * 8),
                "d2h_bytes_per_step": int(4 + logits_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "concurrent_models": conc,
        "graph_replay": graph_leg,
        "config2_pipeline": config2,
        "partitioned": partitioned,
        "clocks": clocks,
        # dominant kernel of the step (61 % of the epoch): the dense contraction, bounded by the tensor pipe
        "roofline": rg,
        # the aggregation (north_star kernel 2).  On this workload its input (49 MB) is L2-resident, so the SURVEY 8(d)
        # algorithmic bytes do not go through HBM: the fraction against the HBM peak is formed from the DRAM traffic ncu
        # measured for this kernel, the fraction against the L2 -> SM fabric from the sectors ncu counted.
        "roofline_aggregation": {"kernel": f"spmm_max_fwd F={f_in} (layer-1 aggregation)", "bound": "l2 (L2 -> SM fabric; input L2-resident)",
                                 "algorithmic_bytes": alg, "avg_ms": spmm_ms, "algorithmic_gb_per_s": achieved,
                                 "edges_per_s": e_prime / (spmm_ms * 1e-3),
                                 "traffic": traffic, "lts_sectors_read": lts,
                                 "traffic_source": "ncu --set full capture committed under profiles/ (per launch)" if traffic else None,
                                 "frac_dram": (traffic / (spmm_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                                 "hbm_peak": hbm_peak, "peak_source": peak_src,
                                 "frac_l2": (lts * 32 / (spmm_ms * 1e-3) / 1e9 / l2_cap) if (lts and l2_cap) else None,
                                 "l2_cap": l2_cap, "l2_cap_unit": "GB/s",
                                 "l2_cap_source": "B300_MICROARCH.md: L2 -> SM throughput cap ~6300 B/clk full chip x the SM clock "
                                                  "sampled in this run",
                                 "note": "the >= 70 % of HBM target is judged on the graph that is NOT L2-resident: "
                                         "`partitioned` (configs[3]) -> aggregation_frac_of_hbm_peak, ncu DRAM bytes for that kernel in "
                                         "profiles/"},
        "gemm_mode": {"default": os.environ.get("PLAGNN_GEMM_PARITY", "2 (parity: two accumulation chains per tile as the two halves of its contraction, weight-gradient "
                                                 "chains of 24 k-blocks; all 19 gradient tensors <= 1e-5 at N = 24 041)"),
                      "fast": fast_gemm},
        "gemm": {"ms_per_step": gemm_ms, "tflops_fp32_equiv": flops_epoch / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                 "flops_per_step": flops_epoch, "backend": os.environ.get("PLAGNN_GEMM", "auto (TMA-fed tcgen05, CTA pairs, 3xTF32)"),
                 "measured_in": "second pass with per-call events (ms_per_step_profiled)"},
        "ms_per_step_profiled": ms_profiled / args.steps,
        "spmm": {"ms_per_step": spmm_all_ms},
        "kernels": kernels[:10],
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_subprocess(args)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bench_kernels_n{world}_s{args.steps}.json"), "w") as fh:
        json.dump(kernels, fh, indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def roofline_gemm(prof, steps, flops_epoch, gemm_ms, bf16_peak, peak_src):
    """Tensor-pipe roofline of the dense contraction (K3): every fp32 product costs three tf32 MMAs (hi*hi, lo*hi, hi*lo), so
    the tensor work issued is 3 x the algorithmic flops; dense tf32 runs at half the bf16 rate, hence peak = measured bf16 / 2."""
    if not gemm_ms:
        return None
    peak = bf16_peak / 2.0
    best = None
    for k, (c, t) in prof.items():
        if k[0] != "gemm" or min(k[1], k[2], k[3]) < 64:
            continue
        tf = 3.0 * 2.0 * k[1] * k[2] * k[3] * c / (t * 1e-3) / 1e12
        if best is None or tf > best[1]:
            best = ("gemm/%d/%d/%d" % (k[1], k[2], k[3]), tf, t / c)
    all_tf = 3.0 * flops_epoch / (gemm_ms * 1e-3) / 1e12
    return {"kernel": "gemm_tma_kernel / gemm_tma_db_kernel (all 23 products of the epoch)", "bound": "tensor", "achieved": all_tf, "peak": peak,
            "unit": "TFLOP/s", "frac": all_tf / peak, "traffic": None,
            "peak_source": peak_src + ": sustained dense bf16 / 2 (tf32 rate)",
            "note": "achieved = 3 tf32 MMAs per fp32 product x algorithmic flops / event time (second, fully instrumented pass)",
            "best_product": None if best is None else {"kernel": best[0], "achieved": best[1], "frac": best[1] / peak,
                                                      "avg_ms": best[2]}}


def dense_flops(n, f_in):
    """Algorithmic dense flops of one epoch: forward + backward, layer-1 input gradient skipped."""
    dims = [f_in, 400, 300, 200]
    fwd = 0
    bwd = 0
    for i in range(3):
        f, o = dims[i], dims[i + 1]
        fwd += 2 * n * (f * f + 2 * f * o)
        bwd += 2 * n * (f * f + 2 * f * o)              # weight gradients
        bwd += 2 * n * f * o                            # d neigh
        if i > 0:
            bwd += 2 * n * (f * o + f * f)              # d x
    for f, o in ((200, 100), (100, 12)):
        fwd += 2 * n * f * o
        bwd += 4 * n * f * o
    return fwd + bwd


# ---------------------------------------------------------------------------------------------------
def oracle_epoch_fn(prob, train_index):
    from oracle import plagnn_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    ids = list(range(prob.num_nodes))
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    go.csc()
    torch.manual_seed(70)
    mo = orc.GNN32Ref(go.ndata["feat"].shape[1], *HIDDEN)
    oo = torch.optim.Adam(mo.parameters(), lr=LR)
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in train_index]

    def step():
        orc.train_epoch(mo, oo, go, go.ndata["feat"], go.ndata["loc"], idx, w)
    return step, orc.num_threads()


def cpu_baseline_subprocess(args):
    """The CPU baseline leg in its own process, so that the GPU arm's process never maps anything under oracle/."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-baseline-only", "--cpu-seconds", str(args.cpu_seconds)]
    if args.nodes:
        cmd += ["--nodes", str(args.nodes)]
    if args.edges:
        cmd += ["--edges", str(args.edges)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "cpu baseline process failed: " + r.stderr[-300:]}


def cpu_baseline(prob, train_index, seconds):
    step, threads = oracle_epoch_fn(prob, train_index)
    step()                                               # warm-up
    t0 = time.perf_counter()
    k = 0
    while k < 3 or (time.perf_counter() - t0 < seconds and k < 200):
        step()
        k += 1
    dt = time.perf_counter() - t0
    return {"value": k / dt, "unit": UNIT, "cores": max(threads, torch.get_num_threads()), "kind": "port",
            "sample": f"{k} full epochs of the same workload (1 warm-up) on the host: DGL-equivalent CPU restatement "
                      f"(torch-CPU GEMMs + C/OpenMP max aggregation), not DGL", "seconds": dt}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    prob, train_index = make_problem(0, args.nodes, args.edges)
    step, threads = oracle_epoch_fn(prob, train_index)
    for _ in range(args.warmup):
        step()
    budget = 150.0
    t0 = time.perf_counter()
    k = 0
    while k < args.steps and (k < 3 or time.perf_counter() - t0 < budget):
        step()
        k += 1
    dt = time.perf_counter() - t0
    n = prob.num_nodes
    val = k / dt
    cores = max(threads, torch.get_num_threads())
    sample = (f"{k} full epochs (of --steps {args.steps}; capped at {budget:.0f} s wall) of the same workload on the host "
              "cores: DGL-equivalent CPU restatement, DGL itself is not installable in this image")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PPI-shaped TSA-state graph (BASELINE configs[1]): N={n}, E={len(prob.ppi_row)} directed + {n} "
                               f"self-loops, F={prob.features.shape[1]}, GNN32 {prob.features.shape[1]}-400-300-200-100-12, full-graph epoch "
                               "(zero_grad, fwd, indexed weighted-BCE, bwd, Adam)",
                   "parallelism": f"host CPU, {cores} threads (the reference's own device for configs[0])",
                   "train_rows": int(len(train_index)), "lr": LR},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    args = parse()
    if args.cpu_baseline_only:
        prob, train_index = make_problem(0, args.nodes, args.edges)
        print(json.dumps(cpu_baseline(prob, train_index, args.cpu_seconds)))
        return
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "scaled":
        from plagnn_b200 import dist_bench
        return dist_bench.run(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
