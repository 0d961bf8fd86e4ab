"""GPU: encoder inputs -- the gather/offset step of encode_subgraphs (gps/models/ss_gnn.py:427-471)
on the device, against a line-by-line torch restatement of the reference code run on the CPU."""
import numpy as np
import pytest
import torch

from helpers import random_batch

pytestmark = pytest.mark.gpu


def reference_encoder_inputs(x, edge_attr, nodes_t, edge_index_t, edge_ptr_t, edge_src_global_t):
    """ss_gnn.py:437-466 verbatim in behaviour (CPU tensors)."""
    num_subgraphs, k = nodes_t.shape
    stacked_nodes = nodes_t.flatten()
    valid_mask = stacked_nodes >= 0
    global_x = x[stacked_nodes.clamp(min=0)]
    if not valid_mask.all():
        global_x = global_x * valid_mask.view(-1, *([1] * (x.dim() - 1))).to(global_x.dtype)
    global_edge_attr = edge_attr[edge_src_global_t] if edge_attr is not None else None
    gei = torch.repeat_interleave(torch.arange(0, num_subgraphs), edge_ptr_t[1:] - edge_ptr_t[:-1]) * k
    gei = gei + edge_index_t
    gb = torch.repeat_interleave(torch.arange(0, num_subgraphs), k)
    return global_x, global_edge_attr, gei, gb


def _batch(seed, G=60, k=4, m=7, **kw):
    from ugs_b200.engine import DeviceBatch
    rng = np.random.default_rng(seed)
    ei, ptr = random_batch(rng, G, 1, 50, deg=2.5, **kw)
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    return rng, ei, ptr, b


@pytest.mark.parametrize("xshape,xdtype", [((5,), torch.float32), ((1,), torch.int64), ((3,), torch.float16),
                                           ((), torch.float32), ((2, 3), torch.float64), ((7,), torch.uint8),
                                           ((32,), torch.float32)])
def test_encoder_inputs_equal_the_reference_ops(cuda_lib, xshape, xdtype):
    from ugs_b200 import encode
    rng, ei, ptr, b = _batch(11, big=[(3, 150)])
    k, m = 4, 7
    (nodes, eidx, eptr, sptr, esrc), _ = b.sample(m, k, "sample", seed=5)
    assert (nodes < 0).any() and (nodes >= 0).any()  # both padded and complete rows are exercised
    N, E = int(ptr[-1]), ei.shape[1]
    g = torch.Generator().manual_seed(1)
    if xdtype.is_floating_point:
        x = torch.randn((N,) + xshape, generator=g, dtype=torch.float32).to(xdtype)
        ea = torch.randn((E, 3), generator=g)
    else:
        x = torch.randint(0, 100, (N,) + xshape, generator=g).to(xdtype)
        ea = torch.randint(0, 4, (E,), generator=g)
    want = reference_encoder_inputs(x, ea, nodes.cpu(), eidx.cpu(), eptr.cpu(), esrc.cpu())
    got = encode.encoder_inputs(x.cuda(), ea.cuda(), nodes, eidx, eptr, esrc)
    for w, gt, name in zip(want, got, ["global_x", "global_edge_attr", "global_edge_index", "global_batch_ptr"]):
        assert gt.dtype == w.dtype and tuple(gt.shape) == tuple(w.shape), name
        assert torch.equal(gt.cpu(), w), name
    # edge_attr=None
    got = encode.encoder_inputs(x.cuda(), None, nodes, eidx, eptr, esrc)
    assert got[1] is None and torch.equal(got[2].cpu(), want[2])


@pytest.mark.parametrize("k,m", [(3, 5), (8, 64), (12, 33)])
def test_fused_path_equals_materialised_path(cuda_lib, k, m):
    """sample_encoder_inputs (UGS_MODE_ENCODER emitted by the sampler) == encoder_inputs(sample(mode='sample'))."""
    from ugs_b200 import encode
    rng, ei, ptr, b = _batch(5, G=80, k=k, m=m, big=[(7, 200), (30, 90)])
    N, E = int(ptr[-1]), ei.shape[1]
    x = torch.randn(N, 6).cuda()
    ea = torch.randn(E, 2).cuda()
    outs, _ = b.sample(m, k, "sample", seed=9)
    a = encode.encoder_inputs(x, ea, outs[0], outs[1], outs[2], outs[4])
    f = encode.sample_encoder_inputs(b, x, ea, m, k, seed=9)
    for u, v in zip(a, f[:4]):
        assert torch.equal(u, v)
    for u, v in zip((outs[0], outs[2], outs[3], outs[4]), f[4]):
        assert torch.equal(u, v)


def test_gather_rows_edge_cases(cuda_lib):
    from ugs_b200 import encode
    src = torch.arange(40, dtype=torch.float32).view(10, 4).cuda()
    idx = torch.tensor([3, -1, 0, 9, -1], dtype=torch.int64).cuda()
    out = encode.gather_rows(src, idx)
    want = torch.stack([src[3], torch.zeros(4).cuda(), src[0], src[9], torch.zeros(4).cuda()])
    assert torch.equal(out, want)
    assert encode.gather_rows(src, idx[:0]).shape == (0, 4)
    with pytest.raises(IndexError):
        encode.gather_rows(src, torch.tensor([10], dtype=torch.int64).cuda(), check=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        encode.gather_rows(src.cpu(), idx.cpu())
    # odd row sizes take the narrower vector paths: 3 bytes, 6 bytes, 12 bytes, 24 bytes
    for width, dt in [(3, torch.uint8), (3, torch.int16), (3, torch.float32), (3, torch.float64)]:
        s = torch.arange(10 * width).view(10, width).to(dt).cuda()
        assert torch.equal(encode.gather_rows(s, idx[[0, 2, 3]]), s[idx[[0, 2, 3]]])
    # a misaligned view (storage offset of one element) still copies correctly
    base = torch.arange(41, dtype=torch.float32).cuda()
    s = base[1:].view(10, 4)
    assert torch.equal(encode.gather_rows(s, idx[[0, 2, 3]]), s[idx[[0, 2, 3]]])


def test_full_size_property(cuda_lib):
    """ZINC-shaped shard at bench size: batch-level ids stay inside their sample's block and x rows are copies."""
    from ugs_b200 import encode, synth
    from ugs_b200.engine import DeviceBatch
    ei, ptr = synth.make_batch("zinc", 16384, seed=3)
    k, m = 8, 64
    b = DeviceBatch(torch.from_numpy(ei).cuda(), torch.from_numpy(ptr).cuda(), k, 0, int(ptr[-1]))
    x = torch.arange(int(ptr[-1]), dtype=torch.int64).cuda().view(-1, 1)
    gx, gea, gei, gb, (nodes, eptr, sptr, esrc), st = encode.sample_encoder_inputs(b, x, None, m, k, seed=1)
    B = nodes.shape[0]
    assert torch.equal(gx.view(-1), nodes.view(-1).clamp(min=0) * (nodes.view(-1) >= 0))
    assert torch.equal(gb, torch.arange(B, device="cuda").repeat_interleave(k))
    owner = torch.repeat_interleave(torch.arange(B, device="cuda"), eptr[1:] - eptr[:-1])
    assert torch.equal(gei[0] // k, owner) and torch.equal(gei[1] // k, owner)
