import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dagl_b200
from oracle import ce_oracle as O
from conftest import load_npz
dev = torch.device("cuda:0")
for head in ["c1_2", "c2_1", "c3_1"]:
    g = load_npz(f"ce_trained_{head}.npz"); p = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    yref, aux = O.ce_forward(p, g["x"], return_aux=True)
    for impl in ("simt", "tc"):
        ce = dagl_b200.CE(in_channels=64, impl=impl); ce.load_state_dict(p); ce = ce.to(dev).eval()
        with torch.no_grad(): y, bits, nnz = ce.forward_debug(g["x"].to(dev))
        inter = {k: v.cpu() for k, v in ce.intermediates(tuple(g["x"].shape)).items()}
        m = O.unpack_mask_bits(bits.cpu(), aux["mask"].shape[-1])
        fl = m != aux["mask"]
        S = aux["S"]; t = aux["mu"].unsqueeze(-1) * aux["gamma"].unsqueeze(-1); be = aux["beta"].unsqueeze(-1)
        margin = ((S - t) + be).abs(); ulp = 1.1920929e-07 * (S.abs() + t.abs() + be.abs())
        qe = (inter["Q"] - aux["Q"]).abs().max() / aux["Q"].abs().max(); ke = (inter["K"] - aux["K"]).abs().max() / aux["K"].abs().max()
        Sg = inter["Q"][0] @ inter["K"][0].t()
        mu_g = inter["Q"][0].double() @ inter["Kbar"][0].double()
        print(head, impl, "flips", int(fl.sum()), "margin/ulp of flips", (margin / ulp)[fl].tolist()[:5], "Qerr %.2e Kerr %.2e" % (qe, ke),
              "mu err/|mu| %.2e" % ((mu_g.float() - aux["mu"][0]).abs().max() / aux["mu"][0].abs().max()),
              "yerr %.2e" % ((y.cpu() - yref).abs().max() / max(yref.abs().max().item(), 1e-30)))
        if fl.any():
            idx = fl.nonzero()[0]; b_, q_, k_ = idx.tolist()
            print("   flip at q", q_, "k", k_, "S_ref", S[b_, q_, k_].item(), "t", t[b_, q_, 0].item(), "beta", be[b_, q_, 0].item(), "margin", ((S - t) + be)[b_, q_, k_].item(), "S from gpu QK (fp32 mm)", Sg[q_, k_].item(), "mu_gpu*gamma", (mu_g[q_].float() * aux["gamma"][0, q_]).item())
