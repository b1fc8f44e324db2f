"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + torch-CPU + the small C/OpenMP kernel in spmm_cpu.c) of the PLA-GNN
message-passing hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the checker or
as the timed CPU baseline.  Nothing under ``pla-gnn_b200/`` imports it.

What it follows in the reference (paths relative to /root/reference):

* ``code/utils.py:41-49``   edge direction (ppi.row -> ppi.col), self-loops appended after the
  COO edges, feature column order [expr | gcn | ecc].
* ``code/model.py:10-31``   GNN32 topology: 3x SAGEConv(.., 'pool') + 2x Linear, leaky_relu(0.01)
  after conv1..3 and liner1, sigmoid after liner2.
* ``code/train.py:89-108``  multi_loss  (class-weighted BCE on probabilities, clamp(1e-9, 10)).
* ``code/train.py:111-126`` weight_cal.
* ``code/train.py:19-40``   protein_loc_correction (label decision).
* ``code/train.py:43-86``   performances_record (AIM / COV / mlACC).
* ``code/train.py:179-180,195-207`` model dims, Adam(lr) defaults, epoch order.
* ``code/main.py:15-29``    scaling (alteration scoring pre-step); ``:32-48`` mat_merge; ``:80-84`` score + ranking.

Parity status
-------------
* loss / class weights / label decision / metrics / scaling: PINNED — checked against the
  reference's own functions, imported from /root/reference/code/{train,main}.py in the build
  container by ``tests/golden/make_golden.py``; the resulting vectors are committed under
  ``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.
* SAGEConv-pool / graph construction: PARITY UNPINNED.  The arithmetic lives in the third-party
  wheel ``dgl_cu113 0.8.2.post1`` (README.md:27), which is neither vendored nor installable
  here.  The restatement follows DGL 0.8's published semantics and is
  made trustworthy by three independent formulations that must agree (dense masked max,
  per-node loop, CSC segment kernel), float64 gradcheck, and structural invariants.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------------------
# C kernel loader
# --------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """Compile oracle/spmm_cpu.c -> oracle/_build/liboracle_spmm.so (idempotent)."""
    out = os.path.join(_HERE, "_build", "liboracle_spmm.so")
    src = os.path.join(_HERE, "spmm_cpu.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, stdout=subprocess.DEVNULL)
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.oracle_spmm_max_f32.argtypes = [i64, i64, p, p, p, i64, p, p, i64]
        lib.oracle_spmm_max_bwd_f32.argtypes = [i64, i64, i64, p, p, i64, p, i64]
        lib.oracle_spmm_sum_f32.argtypes = [i64, i64, p, p, p, p, p, p, i64, p, i64]
        lib.oracle_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().oracle_num_threads())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


# --------------------------------------------------------------------------------------
# a1  graph construction  (utils.py:41-45 + DGL COO->CSC, restated)
# --------------------------------------------------------------------------------------
def add_self_loop(src: np.ndarray, dst: np.ndarray, num_nodes: int):
    """dgl.add_self_loop: append one (i,i) edge per node AFTER the existing edges
    (edge ids E..E+N-1); existing edges, duplicates included, are kept (utils.py:45)."""
    loop = np.arange(num_nodes, dtype=src.dtype)
    return np.concatenate([src, loop]), np.concatenate([dst, loop])


def coo_to_csc(src: np.ndarray, dst: np.ndarray, num_nodes: int):
    """In-edge CSR ("CSC") as DGL materialises it: stable sort of edge ids by destination.
    Returns (indptr[N+1] int32, indices[E] int32 = source of each in-edge, eids[E] int32)."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    perm = np.argsort(dst, kind="stable")
    counts = np.bincount(dst, minlength=num_nodes)
    indptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr.astype(np.int32), src[perm].astype(np.int32), perm.astype(np.int32)


def coo_to_csr(src, dst, num_nodes):
    """Out-edge CSR: same construction with the roles of src and dst swapped."""
    return coo_to_csc(dst, src, num_nodes)


class OracleGraph:
    """Minimal stand-in for the DGLGraph the reference builds in utils.create_graph."""

    def __init__(self, src, dst, num_nodes):
        self.src = np.asarray(src, dtype=np.int64)
        self.dst = np.asarray(dst, dtype=np.int64)
        self.num_nodes = int(num_nodes)
        self.ndata = {}
        self._csc = None
        self._csr = None

    @property
    def num_edges(self):
        return len(self.src)

    def csc(self):
        if self._csc is None:
            self._csc = coo_to_csc(self.src, self.dst, self.num_nodes)
        return self._csc

    def csr(self):
        if self._csr is None:
            self._csr = coo_to_csr(self.src, self.dst, self.num_nodes)
        return self._csr

    def to(self, device):  # CPU only
        return self


def create_graph(ppi, ecc, gcn, loc, expr, uniprot) -> OracleGraph:
    """utils.py:28-51 restated: edges row->col, self-loops, feat=[expr|gcn|ecc] f32, loc f32."""
    n = len(uniprot)
    src, dst = add_self_loop(np.asarray(ppi.row, dtype=np.int64), np.asarray(ppi.col, dtype=np.int64), n)
    g = OracleGraph(src, dst, n)
    g.ndata["loc"] = torch.from_numpy(loc.toarray().astype(np.float32))
    g.ndata["feat"] = torch.tensor(np.hstack((expr, np.hstack((gcn, ecc)))), dtype=torch.float)
    return g


# --------------------------------------------------------------------------------------
# a3/a4  aggregation kernels — three formulations of copy_u + max
# --------------------------------------------------------------------------------------
def spmm_max_c(indptr, indices, x: torch.Tensor):
    """CSC segment kernel (C/OpenMP).  Returns (out, arg) with arg = winning source id."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n = len(indptr) - 1
    out = torch.empty((n, x.shape[1]), dtype=torch.float32)
    arg = torch.empty((n, x.shape[1]), dtype=torch.int32)
    _lib().oracle_spmm_max_f32(n, x.shape[1], _ptr(indptr), _ptr(indices), _ptr(x), x.stride(0),
                               _ptr(out), _ptr(arg), out.stride(0))
    return out, arg


def spmm_max_loop(indptr, indices, x: torch.Tensor):
    """Per-node loop formulation (any float dtype).  np/torch argmax return the FIRST maximum."""
    n = len(indptr) - 1
    out = torch.zeros((n, x.shape[1]), dtype=x.dtype)
    arg = torch.full((n, x.shape[1]), -1, dtype=torch.int32)
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    for v in range(n):
        b, e = int(indptr[v]), int(indptr[v + 1])
        if b == e:
            continue
        seg = x[idx[b:e]]
        val, pos = seg.max(dim=0)
        # torch.max(dim) does not promise the first index on ties; redo ties explicitly
        first = (seg == val.unsqueeze(0)).to(torch.int8).argmax(dim=0)
        out[v] = val
        arg[v] = idx[b:e][first].to(torch.int32)
    return out, arg


def spmm_max_dense(src, dst, num_nodes, x: torch.Tensor):
    """Dense masked-max formulation (small graphs only): values only, no arg."""
    a = torch.zeros((num_nodes, num_nodes), dtype=torch.bool)
    a[torch.as_tensor(dst, dtype=torch.long), torch.as_tensor(src, dtype=torch.long)] = True
    neg = torch.full((), float("-inf"), dtype=x.dtype)
    big = torch.where(a.unsqueeze(-1), x.unsqueeze(0), neg)  # [dst, src, f]
    out = big.max(dim=1).values
    out[~a.any(dim=1)] = 0
    return out


def spmm_max_bwd_c(arg: torch.Tensor, dz: torch.Tensor, n_src: int):
    dz = dz.contiguous()
    dx = torch.empty((n_src, dz.shape[1]), dtype=torch.float32)
    _lib().oracle_spmm_max_bwd_f32(dz.shape[0], n_src, dz.shape[1], _ptr(dz), _ptr(arg), dz.stride(0),
                                   _ptr(dx), dx.stride(0))
    return dx


def spmm_sum_c(indptr, indices, x: torch.Tensor, eids=None, w=None, scale=None):
    """out[v] = scale[v] * sum_{e in in(v)} w[eid(e)] * x[src(e)]  (copy_u/u_mul_e + sum)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n = len(indptr) - 1
    out = torch.empty((n, x.shape[1]), dtype=torch.float32)
    _lib().oracle_spmm_sum_f32(n, x.shape[1], _ptr(indptr), _ptr(indices), _ptr(eids), _ptr(w), _ptr(scale),
                               _ptr(x), x.stride(0), _ptr(out), out.stride(0))
    return out


class _SpMMMax(torch.autograd.Function):
    """DGL GSpMM('copy_lhs','max') forward + its autograd backward (scatter-add by argU)."""

    @staticmethod
    def forward(ctx, x, indptr, indices, use_c):
        if use_c and x.dtype == torch.float32:
            out, arg = spmm_max_c(indptr, indices, x.contiguous())
        else:
            out, arg = spmm_max_loop(indptr, indices, x)
        ctx.save_for_backward(arg)
        ctx.n_src = x.shape[0]
        ctx.use_c = use_c and x.dtype == torch.float32
        return out

    @staticmethod
    def backward(ctx, dz):
        (arg,) = ctx.saved_tensors
        if ctx.use_c:
            return spmm_max_bwd_c(arg, dz, ctx.n_src), None, None, None
        dx = torch.zeros((ctx.n_src, dz.shape[1]), dtype=dz.dtype)
        a = arg.long()
        valid = a >= 0
        cols = torch.arange(dz.shape[1]).expand_as(a)
        dx.index_put_((a[valid], cols[valid]), dz[valid], accumulate=True)
        return dx, None, None, None


def spmm_max(x, indptr, indices, use_c=True):
    return _SpMMMax.apply(x, indptr, indices, use_c)


class _SpMMSum(torch.autograd.Function):
    """copy_u/u_mul_e + sum with optional per-destination scale; backward = transposed SpMM."""

    @staticmethod
    def forward(ctx, x, graph, w, scale):
        indptr, indices, eids = graph.csc()
        ctx.graph, ctx.w, ctx.scale = graph, w, scale
        return spmm_sum_c(indptr, indices, x.contiguous(), eids=eids if w is not None else None, w=w, scale=scale)

    @staticmethod
    def backward(ctx, dz):
        # d x[u] = sum_{e: u->v} w_e * scale[v] * dz[v]  -> out-edge CSR, weights by the same eids
        indptr, indices, eids = ctx.graph.csr()
        d = dz.contiguous()
        if ctx.scale is not None:
            d = (d * ctx.scale.unsqueeze(1)).contiguous()
        dx = spmm_sum_c(indptr, indices, d, eids=eids if ctx.w is not None else None, w=ctx.w, scale=None)
        return dx, None, None, None


def spmm_sum(x, graph, w=None, scale=None):
    return _SpMMSum.apply(x, graph, w, scale)


# --------------------------------------------------------------------------------------
# a2/a3  SAGEConv('pool') and GNN32
# --------------------------------------------------------------------------------------
class SAGEConvPoolRef(nn.Module):
    """dgl.nn.pytorch.SAGEConv(in, out, 'pool') of DGL 0.8.x, restated.

    Parameters: fc_pool Linear(in,in) with bias; fc_self / fc_neigh Linear(in,out) without
    bias; separate ``bias`` Parameter[out] initialised to zero.  Construction order pool, self,
    neigh, then xavier_uniform_(gain=sqrt(2)) on the three weights in that order."""

    def __init__(self, in_feats, out_feats, aggregator_type="pool", feat_drop=0.0, bias=True, norm=None,
                 activation=None, use_c=True):
        super().__init__()
        assert aggregator_type == "pool"
        self.fc_pool = nn.Linear(in_feats, in_feats)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_feats)) if bias else None
        self.use_c = use_c
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, graph: OracleGraph, feat, edge_weight=None):
        assert edge_weight is None
        indptr, indices, _ = graph.csc()
        m = F.relu(self.fc_pool(feat))
        neigh = spmm_max(m, indptr, indices, self.use_c)
        rst = self.fc_self(feat) + self.fc_neigh(neigh)
        if self.bias is not None:
            rst = rst + self.bias
        return rst


class GNN32Ref(nn.Module):
    """model.py:10-31 restated on top of SAGEConvPoolRef (dropout arg accepted and unused)."""

    def __init__(self, in_feats, h1_feats, h2_feats, h3_feats, h4_feats, num_classes, dropout=0.5, use_c=True):
        super().__init__()
        self.conv1 = SAGEConvPoolRef(in_feats, h1_feats, "pool", use_c=use_c)
        self.conv2 = SAGEConvPoolRef(h1_feats, h2_feats, "pool", use_c=use_c)
        self.conv3 = SAGEConvPoolRef(h2_feats, h3_feats, "pool", use_c=use_c)
        self.liner1 = nn.Linear(h3_feats, h4_feats)
        self.liner2 = nn.Linear(h4_feats, num_classes)

    def forward(self, g, in_feat):
        h = F.leaky_relu(self.conv1(g, in_feat))
        h = F.leaky_relu(self.conv2(g, h))
        h = F.leaky_relu(self.conv3(g, h))
        h = F.leaky_relu(self.liner1(h))
        return torch.sigmoid(self.liner2(h))


# --------------------------------------------------------------------------------------
# GNN32 with its discrete decisions handed in (gradient parity without near-tie noise)
# --------------------------------------------------------------------------------------
def forward_with_decisions(model: "GNN32Ref", g: OracleGraph, x: torch.Tensor, dec: dict, tol: float = 1e-5):
    """model.py:19-31 evaluated with every DISCRETE decision of the network taken from `dec` instead of from this
    implementation's own roundings.  GNN32 is piecewise linear up to the final sigmoid; its pieces are selected by
      dec["arg"][l]       int  [N, F_l]  winning source of the max-pool of conv l+1 (argU),
      dec["pool_pos"][l]  bool [N, F_l]  relu'(m[argU]) of that winner (pooled value > 0),
      dec["act"][l]       bool [N, O_l]  leaky_relu branch (output > 0) after conv1..3 (l = 0..2) and liner1 (l = 3).
    Two fp32 implementations that agree to 1e-6 on every activation still pick different pieces wherever two candidates
    are closer than that, and ONE re-routed element moves a weight gradient at the 1e-4 level (the sum it lands in
    cancels).  With the decisions shared, what remains is arithmetic, which is what the 1e-5 gradient bar is about.

    Returns (prob, report).  report[...] counts, per decision kind, the entries where this implementation's own choice
    differs from `dec` and the largest gap of such an entry relative to max|tensor|: the caller asserts that each is a
    near-tie (gap <= tol), i.e. that the shared decision is one this implementation could have taken itself, and that
    every dec["arg"] entry is a real in-edge."""
    indptr, indices, _ = g.csc()
    dst_of = np.repeat(np.arange(g.num_nodes, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    edge_keys = np.unique(dst_of * g.num_nodes + indices.astype(np.int64))
    report = {"arg_diff": 0, "arg_gap": 0.0, "pool_diff": 0, "pool_gap": 0.0, "act_diff": 0, "act_gap": 0.0,
              "entries": 0, "bad_edges": 0}

    def branch(z, mask, slope, kind):
        with torch.no_grad():
            own = z > 0
            diff = own != mask
            report[kind + "_diff"] += int(diff.sum())
            if diff.any():
                report[kind + "_gap"] = max(report[kind + "_gap"], float(z[diff].abs().max() / z.abs().max()))
        return z * torch.where(mask, torch.ones((), dtype=z.dtype), torch.full((), slope, dtype=z.dtype))

    h = x
    for l, conv in enumerate((model.conv1, model.conv2, model.conv3)):
        pre = conv.fc_pool(h)
        a = torch.as_tensor(dec["arg"][l]).long()
        valid = a >= 0
        cols = torch.arange(pre.shape[1]).expand_as(a)
        chosen = pre[a.clamp(min=0), cols]
        with torch.no_grad():
            m_own = F.relu(pre).float().contiguous()
            own_max, own_arg = spmm_max_c(indptr, indices, m_own)
            diff = (own_arg.long() != a) & valid
            report["entries"] += a.numel()
            report["arg_diff"] += int(diff.sum())
            if diff.any():
                gap = (own_max.to(pre.dtype) - F.relu(chosen))[diff]
                report["arg_gap"] = max(report["arg_gap"], float(gap.abs().max() / own_max.abs().max()))
            rows = torch.arange(a.shape[0]).unsqueeze(1).expand_as(a)
            keys = (rows[valid] * g.num_nodes + a[valid]).numpy()
            report["bad_edges"] += int((~np.isin(keys, edge_keys)).sum())
            report["bad_edges"] += int((~valid).sum())          # self-loops: every row has an in-edge (utils.py:45)
        pooled = branch(chosen, torch.as_tensor(dec["pool_pos"][l]) & valid, 0.0, "pool")
        rst = conv.fc_self(h) + conv.fc_neigh(pooled)
        if conv.bias is not None:
            rst = rst + conv.bias
        h = branch(rst, torch.as_tensor(dec["act"][l]), 0.01, "act")
    h = branch(model.liner1(h), torch.as_tensor(dec["act"][3]), 0.01, "act")
    report["near_ties_only"] = (report["bad_edges"] == 0 and report["arg_gap"] <= tol and report["pool_gap"] <= tol
                                and report["act_gap"] <= tol)
    return torch.sigmoid(model.liner2(h)), report


def own_decisions(model: "GNN32Ref", g: OracleGraph, x: torch.Tensor) -> dict:
    """The decisions (see forward_with_decisions) this implementation takes by itself at the precision of `model`."""
    indptr, indices, _ = g.csc()
    dec = {"arg": [], "pool_pos": [], "act": []}
    h = x
    with torch.no_grad():
        for conv in (model.conv1, model.conv2, model.conv3):
            m = F.relu(conv.fc_pool(h))
            neigh, arg = (spmm_max_c(indptr, indices, m.contiguous()) if m.dtype == torch.float32
                          else spmm_max_loop(indptr, indices, m))
            rst = conv.fc_self(h) + conv.fc_neigh(neigh)
            if conv.bias is not None:
                rst = rst + conv.bias
            dec["arg"].append(arg)
            dec["pool_pos"].append(neigh > 0)
            dec["act"].append(rst > 0)
            h = F.leaky_relu(rst)
        dec["act"].append(model.liner1(h) > 0)
    return dec


class GCNSumRef(nn.Module):
    """Synthetic-throughput model family (BASELINE.json configs[3]): L layers of
    h <- act( scale_v * sum_{u->v} w_uv * (h W^T)[u] + b ), leaky_relu between layers, no
    activation after the last.  Not part of the reference; it is the copy_u/u_mul_e + sum
    member of the same kernel family."""

    def __init__(self, dims):
        super().__init__()
        self.lins = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])

    def forward(self, g, x, w=None, scale=None):
        h = x
        for i, lin in enumerate(self.lins):
            h = spmm_sum(F.linear(h, lin.weight), g, w, scale) + lin.bias
            if i + 1 < len(self.lins):
                h = F.leaky_relu(h)
        return h


# --------------------------------------------------------------------------------------
# a6/a7  class weights and loss  (train.py:89-126)
# --------------------------------------------------------------------------------------
def weight_cal(loc_mat: np.ndarray) -> np.ndarray:
    """(labelled_rows - class_count) / class_count, float64[12]  (train.py:111-126)."""
    class_num = loc_mat.sum(axis=0)
    sample_num = int((loc_mat.sum(axis=1) != 0).sum())
    return (sample_num - class_num) / class_num


def multi_loss(inp: torch.Tensor, target: torch.Tensor, i_weight) -> torch.Tensor:
    """Per class i: -(1/R) * sum_r [ t*log(clamp(p,1e-9,10))*w_i + (1-t)*log(clamp(1-p,1e-9,10)) ] / (w_i+1) * 2,
    summed over classes in class order (train.py:100-105).  Keeps the reference's operator
    order so that fp32 rounding matches: (a*w + b) / (w+1) * 2, then sum, then / R."""
    total = 0
    rows = len(inp)
    for i in range(len(i_weight)):
        p = inp[:, i]
        t = target[:, i]
        pos = t * torch.log(torch.clamp(p, 1e-9, 10.0)) * i_weight[i]
        neg = (1 - t) * torch.log(torch.clamp(1 - p, 1e-9, 10.0))
        total = total + (-(((pos + neg) / (i_weight[i] + 1) * 2).sum()) / rows)
    return total


# --------------------------------------------------------------------------------------
# label decision + metrics (train.py:19-86), alteration scaling (main.py:15-29)
# --------------------------------------------------------------------------------------
def protein_loc_correction(loc_proba: torch.Tensor, alpha: float) -> torch.Tensor:
    """Column min-max, row-normalise, per-row threshold max-(max-min)*alpha; returns float64 0/1."""
    mn = loc_proba.min(dim=0).values
    mx = loc_proba.max(dim=0).values
    new = (loc_proba - mn) / (mx - mn)
    new = new / new.sum(dim=1).reshape(-1, 1)
    rmax = new.max(dim=1).values
    rmin = new.min(dim=1).values
    thr = rmax - (rmax - rmin) * alpha
    return (new > thr.unsqueeze(1)).double()


def performances_record(loc_true: torch.Tensor, loc_pred: torch.Tensor):
    """AIM, COV, mlACC averaged over rows (train.py:43-86), vectorised.  The reference sums
    per-row fp32 quotients sequentially; this restatement accumulates the same fp32 quotients
    in float64, so results agree to ~1e-6, not bit-exactly."""
    t = loc_true.detach().cpu().long() == 1
    p = loc_pred.detach().cpu().long() == 1
    inter = (t & p).sum(1).float()
    pred = p.sum(1).float()
    real = t.sum(1).float()
    union = (t | p).sum(1).float()
    aim = torch.where(pred == 0, torch.zeros_like(inter), inter / pred)
    cov = inter / real
    acc = inter / union
    n = len(t)
    return float(aim.double().sum() / n), float(cov.double().sum() / n), float(acc.double().sum() / n)


def scaling(logit_mat: np.ndarray) -> np.ndarray:
    """main.py:15-29: subtract column min, divide by column max, divide rows by their sum."""
    mat = logit_mat - logit_mat.min(0)
    mat = mat / mat.max(0)
    return mat / mat.sum(1, keepdims=True)


def mat_merge(mats) -> np.ndarray:
    """main.py:32-48 on in-memory matrices: float64 accumulator, += scaling(mat) per run, divided by the run count
    (the reference divides by the literal 100, its fixed number of runs)."""
    acc = np.zeros(np.asarray(mats[0]).shape)
    for m in mats:
        acc += scaling(np.asarray(m))
    acc /= len(mats)
    return acc


def alteration_rank(normal_mat: np.ndarray, inter_mat: np.ndarray):
    """main.py:80-84: scaled matrices, relative change, flat indices from the largest to the smallest score.
    The reference uses numpy's default (unstable) argsort and reverses it: NaN first, then descending; the order inside
    a group of equal scores is an artefact of that sort.  Here: stable argsort reversed (ties by descending index)."""
    normal = scaling(np.asarray(normal_mat, dtype=np.float64))
    inter = scaling(np.asarray(inter_mat, dtype=np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        diff = (inter - normal) / normal
    order = np.argsort(diff.reshape(-1), kind="stable")[::-1].copy()
    return normal, inter, diff, order


# --------------------------------------------------------------------------------------
# a9  one epoch, in the order of train.py:195-207
# --------------------------------------------------------------------------------------
def train_epoch(model, optimizer, g, features, labels, train_index, i_weight):
    """zero_grad -> forward -> loss on train rows -> backward -> Adam step.  Returns (logits, loss)."""
    optimizer.zero_grad()
    model.train()
    logits = model(g, features)
    loss = multi_loss(logits[train_index], labels[train_index], i_weight)
    loss.backward()
    optimizer.step()
    return logits, loss
