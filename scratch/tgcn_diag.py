"""Diagnostic: per-parameter gradient error vs the fp64 truth, fused (K7a+K7) vs torch formulation of the dense tail."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tagrec_b200 as T
from tagrec_b200 import tgcn as TG
import test_gpu_parity as P
from helpers import relerr
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
tiny = dict(np.load(os.path.join(G, "tiny.npz"))); tt = dict(np.load(os.path.join(G, "tiny_tgcn.npz")))
truth = dict(np.load(os.path.join(G, "routing_fp64.npz")))

def torch_forward(self, eu, ei, et, ew, u_iw, u_tw, i_uw, i_tw, t_uw, t_iw):
    a_u, a_i, a_t = self.atten1["user"], self.atten1["item"], self.atten1["tag"]
    pj_u, pj_i, pj_t = torch.matmul(eu, a_u.W_2), torch.matmul(ei, a_i.W_2), torch.matmul(et, a_t.W_2)
    eu_iN = a_i.forward(eu, ei, ew, u_iw, pj_i); eu_tN = a_t.forward(eu, et, ew, u_tw, pj_t)
    ei_uN = a_u.forward(ei, eu, ew, i_uw, pj_u); ei_tN = a_t.forward(ei, et, ew, i_tw, pj_t)
    et_uN = a_u.forward(et, eu, ew, t_uw, pj_u); et_iN = a_i.forward(et, ei, ew, t_iw, pj_i)
    mode = os.environ.get("MODE", "torch")
    if mode == "torch_att":      # torch atten2, fused K7 (with torch vec conv)
        zN = self._atten2(torch.cat([eu, ei_uN, et_uN], 0), torch.cat([eu_iN, ei, et_iN], 0), torch.cat([eu_tN, ei_tN, et], 0))
        return torch.split(self._conv_fusion(zN), [eu.shape[0], ei.shape[0], et.shape[0]], dim=0)
    outs = []
    for trip in ((eu, eu_iN, eu_tN), (ei_uN, ei, ei_tN), (et_uN, et_iN, et)):
        z = self._atten2(*trip)
        n = z.shape[0]
        wb = self.conv["bit_level"].weight[:, 0, :, 0]
        bit = torch.relu(torch.einsum('cr,nrd->ncd', wb, z)).reshape(n, -1)
        y = torch.cat([bit, self._vec_conv(z)], 1)
        outs.append(torch.relu(torch.addmm(self.bf, y, self.Wf)))
    return tuple(outs)

def run(label, patch):
    orig = TG.BasicLayer.forward
    if patch: TG.BasicLayer.forward = torch_forward
    model = P._tgcn_model(tiny, tt); model.train()
    lossx = model.loss(torch.tensor(tt["tgcn_batch"], device="cuda"))
    sum(lossx).backward()
    TG.BasicLayer.forward = orig
    rows = []
    for name, p in model.named_parameters():
        want = tt[f"tgcn_grad_{name}"]; t64 = truth[f"tgcn_grad64_{name}"]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(want)
        rows.append((relerr(got, t64) / max(relerr(want, t64), 1e-7), relerr(got, t64), relerr(want, t64), float(np.abs(t64).max()), name))
    rows.sort(reverse=True)
    print("==", label)
    for r in rows[:8]: print("  ratio %.1f err %.2e ref %.2e scale %.2e %s" % r)

run("fused K7a+K7", False)
os.environ["MODE"] = "torch_att"; run("torch atten2 + K7", True)
os.environ["MODE"] = "torch"; run("all torch tail", True)
