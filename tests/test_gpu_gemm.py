"""K3 parity: both GEMM backends vs a float64 torch reference (the tolerance is the north_star's 1e-5
relative, checked norm-wise; the FFMA backend is expected to be ~1e-6, the 3xTF32 tensor-core one too)."""
import pytest
import torch

from plagnn_b200 import ops

pytestmark = pytest.mark.gpu
TOL = 1e-5


def ref_gemm(pairs, bias, act, gate, gate_act):
    acc = 0
    for a, at, b, bt, k in pairs:
        A = a.double().cpu().t() if at else a.double().cpu()
        B = b.double().cpu() if bt else b.double().cpu().t()
        acc = acc + A @ B
    if bias is not None:
        acc = acc + bias.double().cpu()
    if act == ops.ACT_RELU:
        acc = torch.relu(acc)
    elif act == ops.ACT_LEAKY:
        acc = torch.nn.functional.leaky_relu(acc, 0.01)
    elif act == ops.ACT_SIGMOID:
        acc = torch.sigmoid(acc)
    if gate is not None:
        y = gate.double().cpu()
        d = {ops.ACT_RELU: (y > 0).double(), ops.ACT_LEAKY: torch.where(y > 0, 1.0, 0.01),
             ops.ACT_SIGMOID: y * (1 - y)}[gate_act]
        acc = acc * d
    return acc


def rel(a, b):
    return ((a.double().cpu() - b).abs().max() / b.abs().max()).item()


def operand(rows, k, trans, dev, seed, pad):
    g = torch.Generator().manual_seed(seed)
    shape = (k, rows) if trans else (rows, k)
    t = torch.randn(shape, generator=g)
    if pad:
        return ops.aligned(t.to(dev))
    return t.to(dev)        # contiguous: pitch = cols (unaligned when cols % 4 != 0)


BACKENDS = [ops.GEMM_SIMT, ops.GEMM_TCGEN05, ops.GEMM_TMA]
IDS = ["simt", "tcgen05", "tma"]


@pytest.mark.parametrize("backend", BACKENDS, ids=IDS)
@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (300, 200, 100), (1000, 503, 503), (257, 100, 200), (128, 16, 8),
                                   (2500, 400, 12)])
@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_layouts(cuda, backend, m, n, k, at, bt):
    a = operand(m, k, at, cuda, 1, pad=True)
    b = operand(n, k, bt, cuda, 2, pad=True)
    pairs = [(a, at, b, bt, k)]
    got = ops.gemm(m, n, pairs, backend=backend)
    assert rel(got, ref_gemm(pairs, None, 0, None, 0)) < TOL


@pytest.mark.parametrize("backend", BACKENDS, ids=IDS)
def test_gemm_unaligned_operands_and_output(cuda, backend):
    m, n, k = 333, 503, 503
    a = operand(m, k, 0, cuda, 3, pad=False)           # pitch 503: scalar load path
    b = operand(n, k, 0, cuda, 4, pad=False)
    out = torch.empty(m, n, device=cuda)               # pitch 503: scalar store path
    pairs = [(a, 0, b, 0, k)]
    backend0 = backend
    if backend in (ops.GEMM_TCGEN05, ops.GEMM_TMA):
        # the tensor-core path takes 16-byte aligned operands only: explicit request fails loudly, AUTO uses FFMA
        import plagnn_b200 as P
        with pytest.raises(P.PlagnnError):
            ops.gemm(m, n, pairs, out=out, backend=backend)
        backend = ops.GEMM_AUTO
    ops.gemm(m, n, pairs, out=out, backend=backend)
    assert rel(out, ref_gemm(pairs, None, 0, None, 0)) < TOL
    # aligned operands, unaligned output pitch: tensor-core kernel with the scalar store epilogue
    out2 = torch.empty(m, n, device=cuda)
    pairs2 = [(ops.aligned(a), 0, ops.aligned(b), 0, k)]
    ops.gemm(m, n, pairs2, out=out2, backend=backend0)
    assert rel(out2, ref_gemm(pairs, None, 0, None, 0)) < TOL


@pytest.mark.parametrize("backend", BACKENDS, ids=IDS)
@pytest.mark.parametrize("act,gate_act", [(ops.ACT_RELU, 0), (ops.ACT_LEAKY, 0), (ops.ACT_SIGMOID, 0),
                                          (0, ops.ACT_LEAKY), (0, ops.ACT_RELU), (0, ops.ACT_SIGMOID)])
def test_gemm_two_pairs_and_epilogues(cuda, backend, act, gate_act):
    m, n, k1, k2 = 700, 300, 400, 200
    # weights scaled like a layer's (~1/sqrt(k)) so that pre-activations are O(1), as in the network
    a1, b1 = operand(m, k1, 0, cuda, 5, True), operand(n, k1, 0, cuda, 6, True).mul_(k1 ** -0.5)
    a2, b2 = operand(m, k2, 0, cuda, 7, True), operand(n, k2, 0, cuda, 8, True).mul_(k2 ** -0.5)
    bias = torch.randn(n, device=cuda)
    gate = None
    if gate_act:
        gate = ops.aligned(torch.randn(m, n, device=cuda))
        if gate_act == ops.ACT_SIGMOID:
            gate = ops.aligned(torch.sigmoid(gate))
    pairs = [(a1, 0, b1, 0, k1), (a2, 0, b2, 0, k2)]
    got = ops.gemm(m, n, pairs, bias=bias, act=act, gate=gate, gate_act=gate_act, backend=backend)
    want = ref_gemm(pairs, bias, act, gate, gate_act)
    assert rel(got, want) < TOL


@pytest.mark.parametrize("backend", BACKENDS, ids=IDS)
def test_gemm_weight_gradient_shape_split_k(cuda, backend):
    # dW[400 x 503] = dZ^T[400 x 24041] X[24041 x 503]: few output tiles, long K -> split-K path
    nodes, o, f = 24041, 400, 503
    dz = ops.aligned(torch.randn(nodes, o, generator=torch.Generator().manual_seed(1)).to(cuda))
    x = ops.aligned(torch.randn(nodes, f, generator=torch.Generator().manual_seed(2)).to(cuda))
    out = torch.empty(o, f, device=cuda)
    pairs = [(dz, 1, x, 1, nodes)]
    ops.gemm(o, f, pairs, out=out, backend=backend)
    want = ref_gemm(pairs, None, 0, None, 0)
    assert rel(out, want) < TOL
    out2 = torch.empty(o, f, device=cuda)
    ops.gemm(o, f, pairs, out=out2, backend=backend)
    assert torch.equal(out, out2)                      # ordered split-K reduction: bit-stable


# ---- narrow products (one tiny dimension): streaming FFMA kernels, the 12-class head of the network ----
@pytest.mark.parametrize("m,n,k", [(24041, 12, 100), (500, 12, 100), (1000, 16, 300), (777, 5, 37), (64, 1, 1), (65, 9, 129),
                                   (24041, 100, 12), (2500, 400, 12), (333, 503, 7), (100, 2047, 4), (50, 255, 32)])
@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_narrow_layouts(cuda, m, n, k, at, bt):
    a = operand(m, k, at, cuda, 11, pad=True)
    b = operand(n, k, bt, cuda, 12, pad=True)
    pairs = [(a, at, b, bt, k)]
    got = ops.gemm(m, n, pairs, backend=ops.GEMM_NARROW)
    assert rel(got, ref_gemm(pairs, None, 0, None, 0)) < TOL
    auto = ops.gemm(m, n, pairs)                      # AUTO takes the same kernel for these shapes
    assert torch.equal(got, auto)


@pytest.mark.parametrize("n,k1,k2", [(12, 100, 60), (100, 12, 8), (101, 20, 12)])
@pytest.mark.parametrize("act,gate_act", [(ops.ACT_SIGMOID, 0), (ops.ACT_LEAKY, 0), (ops.ACT_RELU, 0), (0, ops.ACT_LEAKY),
                                          (0, ops.ACT_SIGMOID)])
def test_gemm_narrow_two_pairs_and_epilogues(cuda, n, k1, k2, act, gate_act):
    m = 3001
    a1, b1 = operand(m, k1, 0, cuda, 5, True), operand(n, k1, 0, cuda, 6, True).mul_(k1 ** -0.5)
    a2, b2 = operand(m, k2, 0, cuda, 7, True), operand(n, k2, 0, cuda, 8, True).mul_(k2 ** -0.5)
    bias = torch.randn(n, device=cuda)
    gate = None
    if gate_act:
        gate = torch.randn(m, n, device=cuda)
        if gate_act == ops.ACT_SIGMOID:
            gate = torch.sigmoid(gate)
        gate = ops.aligned(gate) if n % 2 == 0 else gate.contiguous()      # odd n: unaligned gate and output pitch
    pairs = [(a1, 0, b1, 0, k1), (a2, 0, b2, 0, k2)]
    out = None if n % 2 == 0 else torch.empty(m, n, device=cuda)
    got = ops.gemm(m, n, pairs, bias=bias, act=act, gate=gate, gate_act=gate_act, out=out, backend=ops.GEMM_NARROW)
    assert rel(got, ref_gemm(pairs, bias, act, gate, gate_act)) < TOL


def test_gemm_narrow_rejects_wide_products(cuda):
    import plagnn_b200 as P
    a, b = operand(500, 100, 0, cuda, 1, True), operand(100, 100, 0, cuda, 2, True)
    with pytest.raises(P.PlagnnError):
        ops.gemm(500, 100, [(a, 0, b, 0, 100)], backend=ops.GEMM_NARROW)


def test_gemm_narrow_wgrad_bias_is_bit_stable(cuda):
    dz = ops.aligned(torch.randn(24041, 12, generator=torch.Generator().manual_seed(5)).to(cuda))
    x = ops.aligned(torch.randn(24041, 100, generator=torch.Generator().manual_seed(6)).to(cuda))
    dw1, db1 = ops.gemm_wgrad_bias(dz, x)
    dw2, db2 = ops.gemm_wgrad_bias(dz, x)
    assert torch.equal(dw1, dw2) and torch.equal(db1, db2)


def test_gemm_tiny_n_falls_to_simt_in_auto(cuda):
    a, b = operand(500, 100, 0, cuda, 1, True), operand(12, 100, 0, cuda, 2, True)
    pairs = [(a, 0, b, 0, 100)]
    got = ops.gemm(500, 12, pairs, act=ops.ACT_SIGMOID)
    assert rel(got, ref_gemm(pairs, None, ops.ACT_SIGMOID, None, 0)) < TOL


def test_tcgen05_and_simt_agree_on_ppi_layer_shape(cuda):
    n, f = 24041, 503
    x = ops.aligned(torch.randn(n, f, generator=torch.Generator().manual_seed(3)).to(cuda))
    w = ops.aligned((torch.randn(f, f, generator=torch.Generator().manual_seed(4)) * 0.05).to(cuda))
    bias = torch.randn(f, device=cuda) * 0.1
    pairs = [(x, 0, w, 0, f)]
    s = ops.gemm(n, f, pairs, bias=bias, act=ops.ACT_RELU, backend=ops.GEMM_SIMT)
    t = ops.gemm(n, f, pairs, bias=bias, act=ops.ACT_RELU, backend=ops.GEMM_TCGEN05)
    assert ((s - t).abs().max() / s.abs().max()).item() < TOL


# ---- TMA-fed kernel: both tile configurations (CTA pair 256 x 256, single CTA 128 x 128) ----
@pytest.fixture(params=["1", "2"], ids=["cg1", "cg2"])
def cta_group(request, monkeypatch):
    monkeypatch.setenv("PLAGNN_TMA_CG", request.param)
    return request.param


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (256, 256, 64), (300, 200, 100), (1000, 503, 503), (257, 100, 200),
                                   (128, 16, 8), (2500, 400, 12), (513, 300, 1000)])
@pytest.mark.parametrize("at,bt", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_tma_layouts_both_tile_configs(cuda, cta_group, m, n, k, at, bt):
    a = operand(m, k, at, cuda, 1, pad=True)
    b = operand(n, k, bt, cuda, 2, pad=True)
    pairs = [(a, at, b, bt, k)]
    got = ops.gemm(m, n, pairs, backend=ops.GEMM_TMA)
    assert rel(got, ref_gemm(pairs, None, 0, None, 0)) < TOL


def test_gemm_tma_two_pairs_epilogues_both_tile_configs(cuda, cta_group):
    m, n, k1, k2 = 700, 300, 400, 200
    a1, b1 = operand(m, k1, 0, cuda, 5, True), operand(n, k1, 0, cuda, 6, True).mul_(k1 ** -0.5)
    a2, b2 = operand(m, k2, 0, cuda, 7, True), operand(n, k2, 0, cuda, 8, True).mul_(k2 ** -0.5)
    bias = torch.randn(n, device=cuda)
    gate = ops.aligned(torch.randn(m, n, device=cuda))
    pairs = [(a1, 0, b1, 0, k1), (a2, 0, b2, 0, k2)]
    got = ops.gemm(m, n, pairs, bias=bias, act=ops.ACT_LEAKY, gate=gate, gate_act=ops.ACT_LEAKY, backend=ops.GEMM_TMA)
    assert rel(got, ref_gemm(pairs, bias, ops.ACT_LEAKY, gate, ops.ACT_LEAKY)) < TOL


def test_gemm_tma_weight_gradient_split_k_both_tile_configs(cuda, cta_group):
    nodes, o, f = 24041, 400, 503
    dz = ops.aligned(torch.randn(nodes, o, generator=torch.Generator().manual_seed(1)).to(cuda))
    x = ops.aligned(torch.randn(nodes, f, generator=torch.Generator().manual_seed(2)).to(cuda))
    out = torch.empty(o, f, device=cuda)
    pairs = [(dz, 1, x, 1, nodes)]
    ops.gemm(o, f, pairs, out=out, backend=ops.GEMM_TMA)
    assert rel(out, ref_gemm(pairs, None, 0, None, 0)) < TOL


@pytest.mark.parametrize("nodes,o,f", [(24041, 400, 503), (24041, 503, 503), (24041, 12, 100), (5000, 300, 400), (3000, 100, 255),
                                       (700, 64, 31), (20000, 700, 600),      # more (tile, split) units than CTA pairs
                                       (1000, 5, 503), (300, 16, 7), (129, 1, 130), (128, 13, 127)])   # narrow kernel (o <= 16)
def test_gemm_wgrad_bias_one_pass(cuda, cta_group, nodes, o, f):
    """dW = dZ^T X with the bias gradient riding along as an extra output column (B column that reads as 1.0)."""
    g1, g2 = torch.Generator().manual_seed(21), torch.Generator().manual_seed(22)
    dz = ops.aligned(torch.randn(nodes, o, generator=g1).to(cuda))
    x = ops.aligned((torch.randn(nodes, f, generator=g2) + 0.5).to(cuda))
    dw, db = ops.gemm_wgrad_bias(dz, x)
    want_w = dz[:, :o].double().cpu().t() @ x[:, :f].double().cpu()
    want_b = dz[:, :o].double().cpu().sum(0)
    assert dw.shape == (o, f) and db.shape == (o,)
    assert rel(dw, want_w) < TOL
    assert ((db.double().cpu() - want_b).abs().max() / want_b.abs().max()).item() < TOL
    # x must be untouched (the ones column is injected in shared memory only)
    assert torch.equal(x, ops.aligned((torch.randn(nodes, f, generator=torch.Generator().manual_seed(22)) + 0.5).to(cuda)))


@pytest.mark.parametrize("persistent", ["1", "0"], ids=["persistent", "one-tile-per-pair"])
def test_gemm_tma_many_tiles_per_cta_pair(cuda, monkeypatch, persistent):
    """More output tiles than CTA pairs: each pair walks several tiles (rings run on across tile boundaries, TMEM is handed
    back by the read-out); same results as the one-tile-per-pair launch, for plain, two-pair + epilogue and split-K products."""
    monkeypatch.setenv("PLAGNN_TMA_PERSISTENT", persistent)
    # plain product with bias + ReLU, 188 tiles
    a, b = operand(24041, 503, 0, cuda, 31, True), operand(503, 503, 0, cuda, 32, True).mul_(503 ** -0.5)
    bias = torch.randn(503, device=cuda)
    pairs = [(a, 0, b, 0, 503)]
    got = ops.gemm(24041, 503, pairs, bias=bias, act=ops.ACT_RELU, backend=ops.GEMM_TMA)
    assert rel(got, ref_gemm(pairs, bias, ops.ACT_RELU, None, 0)) < TOL
    # two pairs, MN-major B (the input-gradient shape without its gate), 188 tiles of 22 k-blocks
    a1, b1 = operand(24041, 400, 0, cuda, 33, True), operand(503, 400, 1, cuda, 34, True).mul_(400 ** -0.5)
    a2, b2 = operand(24041, 300, 0, cuda, 35, True), operand(503, 300, 1, cuda, 36, True).mul_(300 ** -0.5)
    pairs2 = [(a1, 0, b1, 1, 400), (a2, 0, b2, 1, 300)]
    got2 = ops.gemm(24041, 503, pairs2, act=ops.ACT_LEAKY, backend=ops.GEMM_TMA)
    assert rel(got2, ref_gemm(pairs2, None, ops.ACT_LEAKY, None, 0)) < TOL
    # split-K with more (tile, split) units than pairs: 3 x 3 tiles x 44 splits
    a3, b3 = operand(700, 56000, 1, cuda, 37, True), operand(600, 56000, 1, cuda, 38, True)
    pairs3 = [(a3, 1, b3, 1, 56000)]
    got3 = ops.gemm(700, 600, pairs3, backend=ops.GEMM_TMA)
    assert rel(got3, ref_gemm(pairs3, None, 0, None, 0)) < TOL
    # ragged last column tile (n = 300 -> 256 + 64) and a gate (gated launches stay one tile per pair)
    a4, b4 = operand(24041, 200, 0, cuda, 39, True), operand(300, 200, 0, cuda, 40, True).mul_(200 ** -0.5)
    gate = ops.aligned(torch.randn(24041, 300, device=cuda))
    pairs4 = [(a4, 0, b4, 0, 200)]
    got4 = ops.gemm(24041, 300, pairs4, backend=ops.GEMM_TMA)
    assert rel(got4, ref_gemm(pairs4, None, 0, None, 0)) < TOL
    got5 = ops.gemm(24041, 300, pairs4, gate=gate, gate_act=ops.ACT_LEAKY, backend=ops.GEMM_TMA)
    assert rel(got5, ref_gemm(pairs4, None, 0, gate, ops.ACT_LEAKY)) < TOL


# ---- third-generation kernel: 256 x 128 tiles, accumulators double-buffered in tensor memory (gemm_tma_db_kernel) ----------
# Taken by the cost model only for tall products (more tiles than CTA pairs); PLAGNN_TMA_DB_NOW (read per launch) forces it
# on ("1") or off ("0") so that both kernels can be compared on the same operands.
DB_CASES = [  # n, k, pairs, bt, gate, act, bias
    (503, 503, 1, 0, False, "relu", True),      # conv1 fc_pool forward: four full / ragged 128-column tiles
    (400, 503, 2, 0, False, "leaky", True),     # two (A, B) pairs in one accumulation chain; last column tile 16 -> 64 wide
    (300, 200, 1, 1, False, "none", False),     # input gradient: B stored [k x n] (mn-major boxes), column tiles 128 + 128 + 64
    (300, 250, 2, 1, True, "none", False),      # gated input gradient of a SAGE layer: gate images by TMA, one per chunk
    (100, 200, 1, 0, False, "leaky", True),     # one column tile (n_eff = 128)
    (72, 40, 1, 0, False, "sigmoid", True),     # narrow output (n_eff = 128 with 72 real columns), short contraction (2 k-blocks)
    (129, 33, 1, 1, True, "none", False),       # ragged everything: 129 columns (second tile holds one), k = 33
    (96, 24, 1, 0, False, "relu", True),        # one k-block: the two-chain kernel has no second half chain
    (256, 72, 2, 0, True, "leaky", False),      # two pairs of three k-blocks each: the pair switch falls inside a half chain
]


@pytest.mark.parametrize("n,k,pairs,bt,gated,act,with_bias", DB_CASES)
@pytest.mark.parametrize("m", [19201, 24041])
def test_gemm_double_buffered_tiles(cuda, monkeypatch, m, n, k, pairs, bt, gated, act, with_bias):
    act_id = {"none": ops.ACT_NONE, "relu": ops.ACT_RELU, "leaky": ops.ACT_LEAKY, "sigmoid": ops.ACT_SIGMOID}[act]
    ps = []
    for p in range(pairs):
        a = operand(m, k, 0, cuda, 11 + p, pad=True)
        b = operand(n, k, bt, cuda, 21 + p, pad=True)
        ps.append((a, 0, b, bt, k))
    bias = torch.randn(n, generator=torch.Generator().manual_seed(5)).to(cuda) if with_bias else None
    gate = ops.aligned(torch.randn(m, n, generator=torch.Generator().manual_seed(6)).to(cuda)) if gated else None
    gate_act = ops.ACT_LEAKY if gated else ops.ACT_NONE
    ref = ref_gemm(ps, bias, act_id, gate, gate_act)
    res = {}
    # parity mode (the default): the 256 x 128 kernel with two accumulation chains per tile, whatever the cost model says
    monkeypatch.delenv("PLAGNN_GEMM_PARITY", raising=False)
    out = ops.alloc(m, n, cuda, zero=True)
    ops.gemm(m, n, ps, bias=bias, act=act_id, gate=gate, gate_act=gate_act, out=out, backend=ops.GEMM_TMA)
    assert rel(out, ref) < TOL
    two_chain = out.clone()
    # the single-chain kernels (PLAGNN_GEMM_PARITY=0): 256 x 256 tiles and the double-buffered 256 x 128 tiles
    monkeypatch.setenv("PLAGNN_GEMM_PARITY", "0")
    for mode in ("0", "1"):
        monkeypatch.setenv("PLAGNN_TMA_DB_NOW", mode)
        out = ops.alloc(m, n, cuda, zero=True)
        ops.gemm(m, n, ps, bias=bias, act=act_id, gate=gate, gate_act=gate_act, out=out, backend=ops.GEMM_TMA)
        res[mode] = out.clone()
        assert rel(out, ref) < TOL, mode
    # same MMAs in the same order into the same kind of accumulators: the two kernels agree to the bit
    assert torch.equal(res["0"], res["1"])
    # two chains: a different (shorter) accumulation order, so not bit-identical, but the same product
    assert rel(two_chain, res["0"].double().cpu()) < 1e-5
    # nothing written past the row's last 16-byte group (the output is a view of a row-padded buffer; TMA stores clip at the
    # tensor map's extent in 16-byte units, so columns [n, roundup4(n)) of the padding may be written — include/plagnn.h)
    monkeypatch.setenv("PLAGNN_TMA_DB_NOW", "1")
    buf = torch.full((m, ops.pitch_of(n)), 7.0, device=cuda)
    out = buf[:, :n]
    ops.gemm(m, n, ps, bias=bias, act=act_id, gate=gate, gate_act=gate_act, out=out, backend=ops.GEMM_TMA)
    assert torch.equal(out, res["1"]) and bool((buf[:, (n + 3) // 4 * 4:] == 7.0).all())


def test_gemm_double_buffered_ring_variants(cuda, monkeypatch):
    """The ring depths of the double-buffered kernel (raw 4 / lo 3, raw 5 / lo 3, raw 4 / lo 4) give identical results."""
    m, n, k = 19201, 300, 400
    a, b = operand(m, k, 0, cuda, 31, pad=True), operand(n, k, 0, cuda, 32, pad=True)
    monkeypatch.setenv("PLAGNN_GEMM_PARITY", "0")
    monkeypatch.setenv("PLAGNN_TMA_DB_NOW", "1")
    outs = []
    for ring in ("0", "1", "2"):
        monkeypatch.setenv("PLAGNN_TMA_DB_RING", ring)
        outs.append(ops.gemm(m, n, [(a, 0, b, 0, k)], backend=ops.GEMM_TMA).clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel(outs[0], ref_gemm([(a, 0, b, 0, k)], None, 0, None, 0)) < TOL
