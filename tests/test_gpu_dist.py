"""Row-partition choreography on the CUDA kernels (world = 1, 3 pipelined row chunks) against the oracle, and the
row-range aggregation launch it relies on."""
import numpy as np
import pytest
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth
from plagnn_b200.dist import CudaBackend, DistGCN, PartitionedGraph, RowPartitionPlan, dist_gcn_forward_backward
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def test_spmm_row_ranges_tile_the_full_result(cuda):
    n, e, f = 3000, 90000, 96
    sg = synth.scaled_graph(n, e, seed=3, max_degree=4000)
    g = P.graph((sg.src.numpy(), sg.dst.numpy()), num_nodes=n)
    g.chunk = 64
    g = g.to(cuda)
    csc = g.csc()
    assert csc.counts[1] > 0                               # some rows are split over chunks
    x = ops.aligned(torch.randn(n, f, device=cuda))
    w = sg.weight.to(cuda)
    full = ops.spmm_sum(csc, x, w=w)
    out = ops.alloc(n, f, cuda, zero=True)
    for a, b in ((0, 700), (700, 701), (701, 2500), (2500, n)):
        ops.spmm_sum_rows(csc, ops.plan_range(csc, a, b), x, out, w=w)
    assert torch.equal(out, full)


def test_partitioned_gcn_world1_matches_oracle(cuda):
    n, e, f = 2000, 60000, 48
    sg = synth.scaled_graph(n, e, seed=5, max_degree=3000)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    plan = RowPartitionPlan(sg.src, sg.dst, n, 0, 1, chunks=3)
    pg = PartitionedGraph(plan, sg.weight, P.build_csr, cuda)
    model = DistGCN([f, 40, 24], seed=3)
    ref = orc.GCNSumRef([f, 40, 24])
    with torch.no_grad():
        for lin, w, b in zip(ref.lins, model.weights, model.biases):
            lin.weight.copy_(w); lin.bias.copy_(b)
    model = model.to(cuda)
    h0 = ops.alloc(plan.per, f, cuda, zero=True)
    h0[:n].copy_(x)
    mask = (torch.arange(plan.per, device=cuda) < n).float().unsqueeze(1)
    with torch.cuda.device(cuda):
        out, grads = dist_gcn_forward_backward(model, pg, h0, CudaBackend(pg), None, lambda o: o * mask / n)
    go = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
    out_ref = ref(go, x, sg.weight, scale)
    (0.5 * (out_ref ** 2).sum() / n).backward()
    # gathered numbering is chunk-major: map back to node order before comparing
    assert rel_err(out[:n], out_ref) < REL_TOL
    ref_grads = [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]
    for g, gr in zip(grads, ref_grads):
        assert rel_err(g, gr) < 2 * REL_TOL


def _pool_reference(sg, x, n, dims):
    from plagnn_b200.dist import DistSAGEPool
    from tests.test_dist_gloo import _PoolStackRef
    model = DistSAGEPool(dims, seed=3)
    with torch.no_grad():
        for b in model.bias:
            b.add_(0.05)
    ref = _PoolStackRef(model)
    go = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    out = ref(go, x)
    (0.5 * (out ** 2).sum() / n).backward()
    return model, out.detach(), ref.grads()


@pytest.mark.parametrize("mode", ["rows", "cols"])
def test_partitioned_max_pool_world1_matches_oracle(cuda, mode):
    """SAGEConv-pool stack over either partition at world 1 (3 pipelined chunks for the row partition): the choreography
    on the CUDA kernels against the oracle's layers."""
    from plagnn_b200.dist import FeaturePartitionPlan, dist_pool_forward_backward
    n, e, f = 2000, 60000, 48
    dims = [f, 40, 24]
    sg = synth.scaled_graph(n, e, seed=5, max_degree=3000)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    model, out_ref, grads_ref = _pool_reference(sg, x, n, dims)
    plan = (RowPartitionPlan(sg.src, sg.dst, n, 0, 1, chunks=3) if mode == "rows" else FeaturePartitionPlan(sg.src, sg.dst, n, 0, 1))
    pg = PartitionedGraph(plan, None, P.build_csr, cuda, transposed=False)
    model = model.to(cuda)
    h0 = ops.alloc(plan.per, f, cuda, zero=True)
    h0[:n].copy_(x)
    with torch.cuda.device(cuda):
        out, grads = dist_pool_forward_backward(model, pg, h0, CudaBackend(pg), None, lambda o: o / n)
    assert rel_err(out[:n], out_ref) < REL_TOL
    for g, gr in zip(grads, grads_ref):
        assert rel_err(g, gr) < 2 * REL_TOL


def test_feature_partitioned_gcn_world1_matches_oracle(cuda):
    from plagnn_b200.dist import FeaturePartitionPlan
    n, e, f = 2000, 60000, 48
    sg = synth.scaled_graph(n, e, seed=5, max_degree=3000)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    plan = FeaturePartitionPlan(sg.src, sg.dst, n, 0, 1, balance="edges")
    pg = PartitionedGraph(plan, sg.weight, P.build_csr, cuda)
    model = DistGCN([f, 40, 24], seed=3)
    ref = orc.GCNSumRef([f, 40, 24])
    with torch.no_grad():
        for lin, w, b in zip(ref.lins, model.weights, model.biases):
            b.add_(0.05)
            lin.weight.copy_(w); lin.bias.copy_(b)
    model = model.to(cuda)
    h0 = ops.alloc(plan.per, f, cuda, zero=True)
    h0[:n].copy_(x)
    with torch.cuda.device(cuda):
        out, grads = dist_gcn_forward_backward(model, pg, h0, CudaBackend(pg), None, lambda o: o / n)
    go = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
    out_ref = ref(go, x, sg.weight, scale)
    (0.5 * (out_ref ** 2).sum() / n).backward()
    assert rel_err(out[:n], out_ref) < REL_TOL
    for g, gr in zip(grads, [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]):
        assert rel_err(g, gr) < 2 * REL_TOL


@pytest.mark.parametrize("world,rows,feat", [(1, 37, 48), (4, 1000, 256), (8, 513, 96)])
def test_column_block_pack_and_unpack(cuda, world, rows, feat):
    """plagnn_cols_pack / _unpack: x[rows x F] <-> blocks[world][rows][F / world], bit-exact both ways."""
    from plagnn_b200 import _lib
    lib = _lib.load()
    x = ops.alloc(rows, feat, cuda)
    x.copy_(torch.randn(rows, feat, device=cuda))
    fc = feat // world
    blocks = torch.empty((world, rows, fc), device=cuda)
    with torch.cuda.device(cuda):
        _lib.check(lib.plagnn_cols_pack(x.data_ptr(), x.stride(0), rows, feat, world, blocks.data_ptr(), ops._stream()), "pack")
        assert torch.equal(blocks, x.reshape(rows, world, fc).permute(1, 0, 2))
        back = ops.alloc(rows, feat, cuda, zero=True)
        _lib.check(lib.plagnn_cols_unpack(blocks.data_ptr(), rows, feat, world, back.data_ptr(), back.stride(0), ops._stream()), "unpack")
    assert torch.equal(back, x)
    assert lib.plagnn_cols_pack(x.data_ptr(), x.stride(0), rows, feat + 2, world, blocks.data_ptr(), ops._stream()) != 0   # feat % (4 * world)


def _nccl_worker(rank, world, port, out_dir):
    import json
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from plagnn_b200 import dist_bench
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    block = dist_bench.run_partitioned(20000, 600000, 64, 2, 3, rank, world, dev)
    if rank == 0:
        with open(os.path.join(out_dir, "block.json"), "w") as fh:
            json.dump(block, fh)
    dist.destroy_process_group()


def test_two_gpu_nccl_partitions_match_whole_graph_run(cuda, tmp_path):
    """Both partitions, both reducers, on two GPUs through the library's NCCL wrappers: output rows and all-reduced
    gradients of the first step equal the whole-graph run on rank 0 to 1e-5 (a wrong slab offset or block order fails here)."""
    import json
    import os
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29500 + (os.getpid() * 13) % 2000
    mp.start_processes(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    block = json.load(open(tmp_path / "block.json"))
    assert len(block["variants"]) == 4 and block["checks_ok"], block


@pytest.mark.parametrize("f", [32, 64, 20])
def test_source_slab_passes_equal_the_single_pass(cuda, f):
    """plagnn_spmm_sum_slab / _max_slab: the in-edges cut by source range into 3 CSR structures, one pass each, against the
    single pass over the whole structure (sum within fp32 reordering, max values bit-exact, arg equal wherever the maximum is
    attained by one source only)."""
    n, e = 4000, 150000
    sg = synth.scaled_graph(n, e, seed=9, max_degree=5000)
    s, d = sg.src.to(cuda).to(torch.int32), sg.dst.to(cuda).to(torch.int32)
    w = sg.weight.to(cuda)
    full = P.build_csr(d, s, n, False, chunk=128)                      # chunk 128: some rows are split (combine kernel path)
    assert full.counts[1] > 0
    slab_id = s // ((n + 2) // 3)
    slabs, ws = [], []
    for i in range(3):
        m = slab_id == i
        c = P.build_csr(d[m], s[m], n, False, chunk=128)
        slabs.append(c)
        ws.append(w[m][c.eids.long()].contiguous())
    x = ops.alloc(n, f, cuda)
    x.copy_(torch.randn(n, f, device=cuda))
    scale = torch.rand(n, device=cuda) + 0.5
    bias = torch.randn(f, device=cuda)
    ref = ops.spmm_sum(full, x, w=w[full.eids.long()].contiguous(), scale=scale, bias=bias, act=ops.ACT_LEAKY, w_in_csr_order=True)
    got = ops.spmm_sum_slabs(slabs, x, ws=ws, scale=scale, bias=bias, act=ops.ACT_LEAKY)
    assert rel_err(got, ref) < 2e-6
    xr = torch.relu(x)                                                  # ties at 0, like the pooled activations
    ro, ra = ops.spmm_max_fwd(full, ops.aligned(xr))
    go, ga = ops.spmm_max_slabs(slabs, ops.aligned(xr))
    assert torch.equal(go, ro)
    cols = torch.arange(f, device=cuda).expand(n, f)
    has = ra >= 0
    assert torch.equal(ga >= 0, has)
    assert torch.equal(xr[ga.clamp(min=0).long(), cols][has], ro[has])   # the chosen source attains the maximum
    pos = has & (ro > 0)
    assert (ga[pos] != ra[pos]).float().mean().item() < 1e-3            # positive maxima are attained by one source (a.s.)


def test_feature_partition_with_source_slabs_world1(cuda):
    """The feature partition with the slab passes forced on (tiny slab budget) at world 1, both reducers, against the oracle."""
    from plagnn_b200.dist import FeaturePartitionPlan, dist_pool_forward_backward
    n, e, f = 2000, 60000, 48
    sg = synth.scaled_graph(n, e, seed=5, max_degree=3000)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    plan = FeaturePartitionPlan(sg.src, sg.dst, n, 0, 1)
    go = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    h0 = ops.alloc(plan.per, f, cuda, zero=True)
    h0[:n].copy_(x)
    # weighted sum (all widths <= 64 so that every aggregation takes the slab path)
    pg = PartitionedGraph(plan, sg.weight, P.build_csr, cuda, feat=64, slab_mb=0.1)
    assert pg.n_slabs > 2
    model = DistGCN([f, 40, 24], seed=3)
    ref = orc.GCNSumRef([f, 40, 24])
    with torch.no_grad():
        for lin, w_, b_ in zip(ref.lins, model.weights, model.biases):
            lin.weight.copy_(w_); lin.bias.copy_(b_)
    model = model.to(cuda)
    with torch.cuda.device(cuda):
        out, grads = dist_gcn_forward_backward(model, pg, h0, CudaBackend(pg), None, lambda o: o / n)
    scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
    out_ref = ref(go, x, sg.weight, scale)
    (0.5 * (out_ref ** 2).sum() / n).backward()
    assert rel_err(out[:n], out_ref) < REL_TOL
    for g, gr in zip(grads, [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]):
        assert rel_err(g, gr) < 2 * REL_TOL
    # max-pool
    pgm = PartitionedGraph(plan, None, P.build_csr, cuda, transposed=False, feat=64, slab_mb=0.1)
    pmodel, pout_ref, pgrads_ref = _pool_reference(sg, x, n, [f, 40, 24])
    pmodel = pmodel.to(cuda)
    with torch.cuda.device(cuda):
        pout, pgrads = dist_pool_forward_backward(pmodel, pgm, h0, CudaBackend(pgm), None, lambda o: o / n)
    assert rel_err(pout[:n], pout_ref) < REL_TOL
    for g, gr in zip(pgrads, pgrads_ref):
        assert rel_err(g, gr) < 2 * REL_TOL


def _p2p_worker(rank, world, port, out_dir):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from plagnn_b200 import ops
    from plagnn_b200.dist import P2PExchange
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    rows, feat = 1000, 64
    fc = feat // world
    x = ops.alloc(rows, feat, dev)
    x.copy_(torch.arange(rows * feat, device=dev, dtype=torch.float32).reshape(rows, feat) + 1e6 * rank)
    ex = P2PExchange(rows * feat * 4, rank, world, dev)
    ok = True
    for it in range(5):                                   # several rounds: the two window regions alternate
        xc = ex.exchange(x, rows, feat, 0, ops._stream())        # all rows x my columns
        for q in range(world):
            ref = torch.arange(rows * feat, device=dev, dtype=torch.float32).reshape(rows, feat)[:, rank * fc:(rank + 1) * fc] + 1e6 * q
            ok = ok and torch.equal(xc[q * rows:(q + 1) * rows], ref)
        back = ex.exchange(xc.clone(), rows, feat, 1, ops._stream()).clone()      # my rows x all columns again
        ok = ok and torch.equal(back, x)
    ok = ok and ex.error() == 0
    ex.destroy()
    with open(os.path.join(out_dir, f"p2p{rank}.txt"), "w") as fh:
        fh.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_gpu_peer_memory_exchange_round_trip(cuda, tmp_path):
    """plagnn_p2p_*: rows -> columns -> rows between two GPUs over CUDA IPC windows, bit-exact, five rounds."""
    import os
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29500 + (os.getpid() * 17) % 2000
    mp.start_processes(_p2p_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    assert [open(tmp_path / f"p2p{r}.txt").read() for r in range(2)] == ["ok", "ok"]
