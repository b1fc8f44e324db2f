"""K1 parity: device CSR/CSC builder vs the oracle's stable sort — bit-exact integer structure."""
import numpy as np
import pytest
import torch

import plagnn_b200 as P
from oracle import plagnn_oracle as orc
from tests.helpers import random_multigraph

pytestmark = pytest.mark.gpu


def check_csr(src, dst, n, loops, cuda, chunk=128):
    s = torch.as_tensor(src, dtype=torch.int32, device=cuda)
    d = torch.as_tensor(dst, dtype=torch.int32, device=cuda)
    csc = P.build_csr(d, s, n, loops, chunk)
    so, do = (orc.add_self_loop(src.astype(np.int64), dst.astype(np.int64), n) if loops
              else (src.astype(np.int64), dst.astype(np.int64)))
    indptr, indices, eids = orc.coo_to_csc(so, do, n)
    assert np.array_equal(csc.indptr.cpu().numpy(), indptr)
    assert np.array_equal(csc.indices.cpu().numpy(), indices)
    assert np.array_equal(csc.eids.cpu().numpy(), eids)
    # plan: every row covered by ceil(deg/chunk) items (>= 1)
    deg = np.diff(indptr)
    want_items = int(np.maximum(1, -(-deg // chunk)).sum())
    assert csc.counts[0] == want_items
    assert csc.counts[1] == int((deg > chunk).sum())
    return csc


@pytest.mark.parametrize("n,e,loops", [(300, 3000, True), (300, 3000, False), (257, 10, True), (70000, 500000, True),
                                       (1, 0, True), (5, 0, False), (1000, 200000, False)])
def test_csr_build_bit_exact(cuda, n, e, loops):
    src, dst = random_multigraph(n, e, n + e) if e else (np.zeros(0, np.int32), np.zeros(0, np.int32))
    check_csr(src, dst, n, loops, cuda)


def test_csr_build_hubs_empty_rows_and_duplicates(cuda):
    src, dst = random_multigraph(5000, 40000, 3, hubs=4, hub_deg=3000, isolated=50)
    src[:100], dst[:100] = 7, 60          # 100 duplicate edges
    csc = check_csr(src, dst, 5000, False, cuda)
    assert csc.counts[1] >= 4
    check_csr(src, dst, 5000, True, cuda, chunk=32)


def test_csr_and_csc_are_transposes(cuda):
    src, dst = random_multigraph(400, 5000, 11)
    g = P.graph((src, dst), num_nodes=400).add_self_loop().to(cuda)
    csc, csr = g.csc(), g.csr()
    a = torch.zeros(400, 400)
    a.index_put_((csc.indices.cpu().long(), torch.repeat_interleave(torch.arange(400), csc.degrees.cpu().long())),
                 torch.ones(csc.num_edges), accumulate=True)          # a[src, dst]
    b = torch.zeros(400, 400)
    b.index_put_((torch.repeat_interleave(torch.arange(400), csr.degrees.cpu().long()), csr.indices.cpu().long()),
                 torch.ones(csr.num_edges), accumulate=True)
    assert torch.equal(a, b)


def test_csr_build_rejects_out_of_range_ids(cuda):
    s = torch.tensor([0, 5], dtype=torch.int32, device=cuda)
    d = torch.tensor([1, 2], dtype=torch.int32, device=cuda)
    with pytest.raises(P.PlagnnError):
        P.build_csr(d, s, 3, False)


def test_full_size_ppi_structure(cuda):
    """BASELINE configs[1] size: N = 24 041, E = 1.4 M.  Size-independent properties + oracle equality."""
    from plagnn_b200 import synth
    prob = synth.ppi_problem(state="inter")
    csc = check_csr(prob.ppi_row, prob.ppi_col, prob.num_nodes, True, cuda)
    ip = csc.indptr.cpu().numpy()
    assert (np.diff(ip) >= 1).all() and ip[-1] == 1_400_000 + 24041
    last = csc.indices.cpu().numpy()[ip[1:] - 1]
    assert np.array_equal(last, np.arange(24041))     # the self-loop closes every row
