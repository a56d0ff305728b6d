"""CPU restatement of the reference ResBlock chain -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this module; the product (``dagl_b200``) never
does.  Pinned against the UNMODIFIED reference class (``common.ResBlock``, DN_Gray/model/common.py:59-79) by
tests/test_oracle.py::test_resblock_oracle_matches_reference (live import, ``torch.equal``) and by the committed fixture
tests/golden/resblock_chain.npz (made by oracle/make_golden_resblock.py).
"""
import torch
import torch.nn.functional as F


def init_resblock_params(seed: int, prelu_n: int = 1, bias: bool = True):
    """Parameters of one ResBlock in the reference's state_dict layout (body.0 / body.1 / body.2), nn.Conv2d-style
    uniform init (bound 1/sqrt(fan_in)), PReLU slope(s) around the default 0.25."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / (64 * 9) ** 0.5
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * bound
    p = {"body.0.weight": u(64, 64, 3, 3), "body.2.weight": u(64, 64, 3, 3),
         "body.1.weight": 0.25 + 0.2 * (torch.rand(prelu_n, generator=g) - 0.5)}
    if bias:
        p["body.0.bias"] = u(64)
        p["body.2.bias"] = u(64)
    return p


def resblock_forward(p, x: torch.Tensor, res_scale: float = 1.0) -> torch.Tensor:
    """common.py:75-79: ``res = self.body(x).mul(self.res_scale); res += x`` with body = conv, PReLU, conv
    (common.py:66-71; conv = default_conv, common.py:8-11: Conv2d(k=3, padding=1))."""
    h = F.conv2d(x, p["body.0.weight"], p.get("body.0.bias"), padding=1)
    h = F.prelu(h, p["body.1.weight"])
    h = F.conv2d(h, p["body.2.weight"], p.get("body.2.bias"), padding=1)
    res = h.mul(res_scale)
    res += x
    return res


def chain_forward(blocks, x: torch.Tensor, res_scale: float = 1.0) -> torch.Tensor:
    """nn.Sequential of ResBlocks (CES.RBS1 / RBS2, dagl.py:86-101; RR.body, dagl.py:27-34)."""
    for p in blocks:
        x = resblock_forward(p, x, res_scale)
    return x
