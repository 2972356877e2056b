"""Debug driver: TC training step vs the fp64 oracle, prints per-tensor relative L2 errors."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from oracle import nic_oracle as O  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
case = sys.argv[2] if len(sys.argv) > 2 else "mip2"
dev = torch.device("cuda:0")
T = lambda a, dt=None: (torch.as_tensor(np.ascontiguousarray(a)).to(dt) if dt else torch.as_tensor(np.ascontiguousarray(a))).to(dev)
size = 256
grids = I.make_grids(size, 2, seed=60)
params = I.make_mlp(73, seed=61, gain=1.5)
rng = np.random.default_rng(64)
if case == "mip2":
    mip, fl, nc, crop = 2, 0, 3, 64
    coord = np.array([[0, 0], [0, 0], [0, 0]])
elif case == "mip0":
    mip, fl, nc, crop = 0, 0, 2, 128
    coord = rng.integers(0, size - crop + 1, (nc, 2))
else:
    mip, fl, nc, crop = 5, 1, 5, 5
    coord = rng.integers(0, (size >> mip) - crop + 1, (nc, 2))
img = I.box_mips(I.make_image(size, 2, seed=62), 8)[mip]
target = np.concatenate([img[:, c[0]:c[0] + crop, c[1]:c[1] + crop].reshape(3, -1).T for c in coord], 0)
noise = I.make_noise(nc * crop * crop, 73, 8, 63) if "nonoise" not in sys.argv else None
loss, out, grads, dg0, dg1 = O.train_forward_backward(grids, params, coord, target, fl, mip, 1, noise, size=crop)
fp = [T(a) for a in grids]
pt = [T(p) for p in params]
m = L.make_mlp(pt)
g = [torch.zeros_like(p) for p in pt]
gm = L.make_mlp_grad(g)
g0t, g1t = fp[2 * fl], fp[2 * fl + 1]
d0, d1 = torch.zeros_like(g0t), torch.zeros_like(g1t)
ls = torch.zeros(4, device=dev)
o = torch.empty((nc * crop * crop, 3), device=dev)
geom = L.make_geom(L.METHOD_2D, g0t, g1t, crop, nc, mip - 2 * (fl + 1), mip, 6, L.PE_TRIANGULAR)
h = L.handle(dev)
coord_t, target_t = T(coord, torch.int64), T(target)
noise_t = T(noise) if noise is not None else None
L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(g0t), L.ptr(g1t), L.ptr(coord_t), C.byref(m),
                                           L.ptr(target_t), L.ptr(noise_t), 0, 0, 0, 0, C.byref(gm), L.ptr(d0), L.ptr(d1),
                                           L.ptr(ls), L.ptr(o), L.PRECISIONS[prec], L.stream_ptr(dev)))
torch.cuda.synchronize()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))
print("loss", float(ls[0]) / (nc * crop * crop * 3), loss, "out max err", np.abs(o.cpu().numpy() - out).max())
for t, k in zip(g, ("W1", "b1", "W2", "b2", "W3", "b3")):
    a = t.cpu().numpy()
    print(k, "rel", rel(a, grads[k]), "ratio", float((a * grads[k]).sum() / (grads[k] ** 2).sum()))
w1 = g[0].cpu().numpy()
for lo, hi, name in ((0, 48, "g0cols"), (48, 60, "g1cols"), (60, 72, "pecols"), (72, 73, "lod")):
    print(" W1", name, rel(w1[:, lo:hi], grads["W1"][:, lo:hi]))
print("dG0 rel", rel(d0.cpu().numpy(), dg0), "ratio", float((d0.cpu().numpy() * dg0).sum() / (dg0 ** 2).sum()))
print("dG1 rel", rel(d1.cpu().numpy(), dg1), "ratio", float((d1.cpu().numpy() * dg1).sum() / (dg1 ** 2).sum()))
b1 = g[1].cpu().numpy()
print("b1 ratio per j:", np.round(b1 / grads["b1"], 2))
w1 = g[0].cpu().numpy()
print("W1 row rel:", np.round([rel(w1[j], grads["W1"][j]) for j in range(64)], 2))
print("W1 col rel:", np.round([rel(w1[:, c], grads["W1"][:, c]) for c in range(73)], 2))
