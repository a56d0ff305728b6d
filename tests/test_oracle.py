"""CPU tests: pin the oracle (oracle/ce_oracle.py) to the golden fixtures that
oracle/make_golden.py produced by running the UNMODIFIED reference, and — when
the reference is mounted (build container only) — to the live reference."""
import os

import pytest
import torch

from conftest import have_reference, import_reference, load_npz
from oracle import ce_oracle as O
from oracle import ref_loader as R


def test_same_pad_rule():
    # dagl.py:126-136; SURVEY §8 a1 table
    assert O.same_pad_amounts(64, 7, 4) == (1, 2)      # n % 4 == 0
    assert O.same_pad_amounts(65, 7, 4) == (3, 3)      # n % 4 == 1
    assert O.same_pad_amounts(66, 7, 4) == (2, 3)      # n % 4 == 2
    assert O.same_pad_amounts(67, 7, 4) == (2, 2)      # n % 4 == 3
    for n in (1, 7, 30, 41, 256):
        assert O.same_pad_amounts(n, 7, 1) == (3, 3)
    assert O.num_queries(256, 256) == (64, 64)
    assert O.num_queries(30, 41) == (8, 11)


def test_cfg1_bit_exact(rand_weights):
    g = load_npz("ce_cfg1_64x64.npz")
    y, aux = O.ce_forward(rand_weights, g["x"], return_aux=True)
    assert torch.equal(y, g["y"])
    assert torch.equal(O.pack_mask_bits(aux["mask"]), g["mask_bits"])
    assert torch.equal(aux["mask"].sum(-1).to(torch.int32), g["nnz"])
    assert torch.equal(aux["gamma"], g["gamma"]) and torch.equal(aux["beta"], g["beta"])
    assert torch.equal(aux["mu"], g["mu"])
    assert torch.equal(aux["Q"][0, :8], g["Q_head"]) and torch.equal(aux["K"][0, :8], g["K_head"])
    assert torch.equal(aux["S"][0, :8, :64], g["S_head"])


def test_ragged_shapes(rand_weights):
    g = load_npz("ce_ragged.npz")
    for i in range(4):
        x, yref = g[f"x{i}"], g[f"y{i}"]
        y, aux = O.ce_forward(rand_weights, x, return_aux=True)
        assert torch.equal(y, yref), f"shape {tuple(x.shape)}"
        assert torch.equal(aux["mask"].sum(-1).to(torch.int32), g[f"nnz{i}"])
        if f"mask_bits{i}" in g:
            assert torch.equal(O.pack_mask_bits(aux["mask"]), g[f"mask_bits{i}"])


def test_reference_smoke_shape(rand_weights):
    g = load_npz("ce_demo_smoke.npz")                    # Demosaic/model/dagl.py:280-284
    y = O.ce_forward(rand_weights, g["x"])
    assert y.shape == (2, 16, 16, 16)
    assert torch.equal(y, g["y"])


@pytest.mark.parametrize("head", ["c1_2", "c2_1", "c3_1", "c3_3"])
def test_trained_heads(head):
    g = load_npz(f"ce_trained_{head}.npz")
    p = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    y, aux = O.ce_forward(p, g["x"], return_aux=True)
    assert torch.equal(y, g["y"])
    assert torch.equal(O.pack_mask_bits(aux["mask"]), g["mask_bits"])
    if head == "c3_3":                                   # whole head masked out -> exactly zero
        assert g["nnz"].sum() == 0 and y.abs().max() == 0


def test_chunked_matches_reference_order(rand_weights):
    g = load_npz("ce_cfg1_64x64.npz")
    y, nnz = O.ce_forward_chunked(rand_weights, g["x"], chunk=37, return_nnz=True)
    assert (y - g["y"]).abs().max() <= 2e-6 * g["y"].abs().max()
    assert torch.equal(nnz.to(torch.int32), g["nnz"])


def test_fp64_arbitration(rand_weights):
    """fp64 evaluation of the same restatement: the fp32 reference sits within 1e-4 of it
    and the neighbour masks agree (0 flips at 64^2)."""
    g = load_npz("ce_cfg1_64x64.npz")
    p64 = {k: v.double() for k, v in rand_weights.items()}
    y64, aux64 = O.ce_forward(p64, g["x"].double(), return_aux=True)
    assert (y64.float() - g["y"]).abs().max() <= 1e-4 * g["y"].abs().max()
    flips = (O.pack_mask_bits(aux64["mask"]) != g["mask_bits"]).sum().item()
    assert flips == 0


def test_mask_bit_packing_roundtrip():
    gen = torch.Generator().manual_seed(3)
    m = torch.rand(2, 5, 77, generator=gen) > 0.5
    w = O.pack_mask_bits(m)
    assert w.shape == (2, 5, 3) and w.dtype == torch.int32
    assert torch.equal(O.unpack_mask_bits(w, 77), m)


def test_algorithmic_work_figures():
    # SURVEY §8(d) / Appendix D
    assert abs(O.algorithmic_flops(1, 256, 256) / 1e9 - 526.1) < 0.1
    assert abs(O.algorithmic_bytes(1, 256, 256) / 1e6 - 63.0) < 0.1
    assert abs(O.algorithmic_flops(1, 64, 64) / 1e9 - 2.055) < 0.001


# ---------------------------------------------------------------------------
# live reference (build container only; /root/reference is absent on the GPU box)
# ---------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not have_reference(), reason="/root/reference not mounted")


@needs_ref
@pytest.mark.parametrize("shape", [(1, 64, 20, 24), (2, 64, 17, 13), (1, 64, 9, 33)])
def test_live_reference_ce(shape):
    ref = import_reference()
    torch.manual_seed(11)
    ce = ref.CE(in_channels=64).eval()
    x = torch.randn(*shape)
    with torch.no_grad():
        yref = ce(x)
    y = O.ce_forward(dict(ce.state_dict()), x)
    assert torch.equal(y, yref)


@needs_ref
def test_live_reference_ces():
    ref = import_reference()
    torch.manual_seed(5)
    ces = ref.CES(in_channels=64).eval()
    x = torch.randn(1, 64, 16, 20)
    with torch.no_grad():
        yref = ces(x)
    y = O.ces_forward(dict(ces.state_dict()), x)
    assert (y - yref).abs().max() <= 1e-5 * yref.abs().max()


@needs_ref
def test_state_dict_compat_with_reference():
    """dagl_b200.CE / CES carry the reference's parameter names and shapes, so reference
    checkpoints load unchanged (SURVEY §5 checkpoint row, §8b)."""
    import dagl_b200
    ref = import_reference()
    rce, mce = ref.CE(in_channels=64), dagl_b200.CE(in_channels=64)
    assert {k: tuple(v.shape) for k, v in rce.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in mce.state_dict().items()}
    mce.load_state_dict(rce.state_dict(), strict=True)
    rces, mces = ref.CES(in_channels=64), dagl_b200.CES(in_channels=64)
    assert {k: tuple(v.shape) for k, v in rces.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in mces.state_dict().items()}
    mces.load_state_dict(rces.state_dict(), strict=True)


@needs_ref
def test_patch_reference_swaps_all_heads_and_shares_parameters():
    import types
    import dagl_b200
    ref = import_reference()
    args = types.SimpleNamespace(n_resblocks=16, n_feats=64, n_colors=1, res_scale=1, rgb_range=1.0)
    net = ref.RR(args)
    before = dict(net.named_parameters())
    keys_before = list(net.state_dict().keys())
    n = dagl_b200.patch_reference(net)
    assert n == 12
    assert list(net.state_dict().keys()) == keys_before
    after = dict(net.named_parameters())
    assert all(after[k] is before[k] for k in before)
    assert all(isinstance(getattr(net.body[8], f"c{s}_{h}"), dagl_b200.CE) for s in (1, 2, 3) for h in (1, 2, 3, 4))


def test_bench_workload_generator_matches_oracle_init():
    """bench.py carries its own seeded head generator (the product arm must not import oracle/); it has to produce the
    weights the oracle-side generator produces, so that both bench arms and the tests talk about the same head."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    params, x = bench.workload_tensors(3)
    ref = O.init_ce_params(1003, in_channels=64)
    assert set(params) == set(ref)
    assert all(torch.equal(params[k], ref[k]) for k in ref)
    assert x.shape == (1, 64, 256, 256)


def test_topk_oracle_matches_the_legacy_reference_class():
    """oracle.ce_forward_topk against outputs of the reference's own leftover fixed-top-k class
    (GReccR2b_3mh_1-checkpoint.py, run by oracle/make_golden_topk.py).  The golden is ``(b + W y)[:, :16] - b[:, :16]`` with a
    selector W, so it carries one fp32 rounding of |b| + |y|: compared to 1e-5 of the output range, not bit for bit."""
    g = load_npz("ce_topk_legacy.npz")
    w = load_npz("ce_rand_w.npz")
    for tag in ("a", "b", "c"):
        y = O.ce_forward_topk(w, g[f"x_{tag}"], int(g[f"k_{tag}"]))
        ref = g[f"y_{tag}"]
        assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), tag


# ---- ResBlock chain oracle (common.py:59-79) ---------------------------------------------------------------------
def _resblock_golden():
    import numpy as np
    from oracle import resblock_oracle as RB
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "resblock_chain.npz"))
    for tag in ("a", "b"):
        nb = int(z[f"{tag}_nb"])
        blocks = [RB.init_resblock_params(100 * ord(tag) + i) for i in range(nb)]
        wsum = sum(float(v.double().abs().sum()) for p in blocks for v in p.values())
        assert abs(wsum - float(z[f"{tag}_wsum"])) <= 1e-9 * wsum, "seeded weights differ from the ones the fixture was made with"
        yield tag, blocks, torch.from_numpy(z[f"{tag}_x"]), torch.from_numpy(z[f"{tag}_y"]), float(z[f"{tag}_res_scale"])


def test_resblock_oracle_matches_golden():
    """oracle/resblock_oracle.py reproduces the outputs the UNMODIFIED reference ResBlock chain produced (bit-exact)."""
    from oracle import resblock_oracle as RB
    n = 0
    for tag, blocks, x, y, rs in _resblock_golden():
        assert torch.equal(RB.chain_forward(blocks, x, rs), y), tag
        n += 1
    assert n == 2


@pytest.mark.skipif(not R.available("DN_Gray"), reason="reference sources not present")
def test_resblock_oracle_matches_reference():
    """Live: the reference's own common.ResBlock (built as CES builds it, dagl.py:86-101) vs the oracle, torch.equal."""
    import torch.nn as nn
    from oracle import resblock_oracle as RB
    common = R.load_task("DN_Gray").common
    for prelu_n, rs in ((1, 1), (1, 0.1)):
        blocks = [RB.init_resblock_params(7 + i, prelu_n) for i in range(2)]
        seq = nn.Sequential(*[common.ResBlock(common.default_conv, n_feats=64, kernel_size=3, act=nn.PReLU(prelu_n), res_scale=rs)
                              for _ in blocks]).eval()
        for blk, p in zip(seq, blocks):
            blk.load_state_dict(p)
        x = torch.randn(1, 64, 11, 13, generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            assert torch.equal(seq(x), RB.chain_forward(blocks, x, rs))


@pytest.mark.skipif(not R.available("DN_Gray"), reason="reference sources not present")
def test_patch_resblocks_finds_every_reference_resblock_and_keeps_cpu_path():
    """patch_reference(..., fuse_resblocks=True) rebinds the nn.Sequential containers that hold ResBlocks (RR.body: 16,
    CES.RBS1 / RBS2: 4 + 4); with a CPU input the fused containers run the reference's own modules (bit-identical)."""
    import dagl_b200
    from dagl_b200 import resblock as DR
    ref = R.load_task("DN_Gray")
    torch.manual_seed(0)
    net = ref.dagl.RR(R.rr_args("DN_Gray")).eval()
    x = torch.randn(1, 64, 8, 8)
    with torch.no_grad():
        want = net.body[0](x).clone()
        assert DR.patch_resblocks(net) == 24
        assert sum(1 for m in net.modules() if DR.is_resblock(m)) == 24
        got = net.body[0](x)
        ces = net.body[8]
        assert type(ces).__name__ == "CES"
        y = x
        for blk in ces.RBS1:                                   # the container's own modules, one by one
            y = blk(y)
        assert torch.equal(ces.RBS1(x), y)                     # the chained Sequential.forward on a CPU input
        # the overrides are class-level, so nn.DataParallel's replication (reference wrapper, model/__init__.py:101-103)
        # keeps them AND binds them to the replica's own sub-modules
        dagl_b200.patch_reference(net, fuse_resblocks=True)
        rep_ces = ces._replicate_for_data_parallel()
        rep_seq = ces.RBS1._replicate_for_data_parallel()
        assert type(rep_ces).forward is type(ces).forward and getattr(type(rep_ces), "_dagl_fused_stages", False)
        assert DR.is_fused(rep_seq) and "forward" not in rep_seq.__dict__ and "forward" not in rep_ces.__dict__
        assert type(ces).__name__ == "CES" and isinstance(ces, ref.dagl.CES)
        n_before = len(net.state_dict())
        dagl_b200.patch_reference(net)                         # idempotent: no second subclass layer
        assert type(ces).__mro__[1] is ref.dagl.CES and len(net.state_dict()) == n_before
    assert torch.equal(got, want)
