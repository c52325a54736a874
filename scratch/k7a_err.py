import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tagrec_b200 as T
from tagrec_b200.tgcn import BasicLayer, TgcnMixFn
from helpers import relerr
dev = "cuda"
n = 333
g = torch.Generator().manual_seed(5)
layer = BasicLayer(64, 64, 32, 10, 32, 8).to(dev)
with torch.no_grad():
    for p in layer.parameters():
        p.copy_(torch.randn(p.shape, generator=g) * 0.2)
xs = [torch.randn(n, 64, generator=g) for _ in range(3)]
up_z, up_f = torch.randn(n, 3, 64, generator=g), torch.randn(n, 48, generator=g)
par = (layer.U, layer.q.reshape(-1), layer.p.reshape(-1)) + tuple(m.weight.reshape(m.weight.shape[0], -1) for m in layer.conv["vec_level"].values())

def torch_path(xs_, P, dt):
    uit = torch.stack(xs_, dim=1)
    a = torch.relu(uit @ P["U"] + P["q"]) @ P["p"].T
    zr = torch.softmax(a, dim=1) * uit
    x = zr.unsqueeze(1)
    vec = []
    for j in range(1, 4):
        w = P[f"conv.vec_level.conv_{j}.weight"]
        w2 = w.reshape(w.shape[0], -1)
        pos = [zr[:, p:p + j, :].reshape(zr.shape[0], -1) @ w2.t() for p in range(4 - j)]
        vec.append(torch.relu(torch.stack(pos, dim=2)).reshape(zr.shape[0], -1))
    return zr, torch.cat(vec, -1)

# fp64 truth
P64 = {k: v.detach().cpu().double().requires_grad_(True) for k, v in layer.named_parameters()}
r64 = [t.clone().double().requires_grad_(True) for t in xs]
z64, f64 = torch_path(r64, P64, torch.float64)
((z64 * up_z.double()).sum() + (f64 * up_f.double()).sum()).backward()
# fused
lv = [t.clone().to(dev).requires_grad_(True) for t in xs]
z, xf = TgcnMixFn.apply(*lv, *par)
((z * up_z.to(dev)).sum() + (xf * up_f.to(dev)).sum()).backward()
fused = {"z": z, "xf": xf, "gx0": lv[0].grad, "gx1": lv[1].grad, "gx2": lv[2].grad}
named = dict(layer.named_parameters())
for k in ("U", "q", "p"): fused["g" + k] = named[k].grad.clone()
layer.zero_grad()
# torch fp32 on GPU
P32 = dict(layer.named_parameters())
l2 = [t.clone().to(dev).requires_grad_(True) for t in xs]
z2, f2 = torch_path(l2, P32, torch.float32)
((z2 * up_z.to(dev)).sum() + (f2 * up_f.to(dev)).sum()).backward()
tor = {"z": z2, "xf": f2, "gx0": l2[0].grad, "gx1": l2[1].grad, "gx2": l2[2].grad}
for k in ("U", "q", "p"): tor["g" + k] = named[k].grad.clone()
truth = {"z": z64, "xf": f64, "gx0": r64[0].grad, "gx1": r64[1].grad, "gx2": r64[2].grad, "gU": P64["U"].grad, "gq": P64["q"].grad.reshape(-1), "gp": P64["p"].grad.reshape(-1)}
for k in truth:
    t = truth[k].detach().numpy()
    print("%4s fused %.2e  torch32 %.2e" % (k, relerr(fused[k].detach().cpu().numpy().reshape(t.shape), t), relerr(tor[k].detach().cpu().numpy().reshape(t.shape), t)))
