"""Debug: single-sample gradient contributions, TC path vs fp64 oracle (N = 128: one tile)."""
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
dev = torch.device("cuda:0")
T = lambda a, dt=None: (torch.as_tensor(np.ascontiguousarray(a)).to(dt) if dt else torch.as_tensor(np.ascontiguousarray(a))).to(dev)
size = 256
grids = I.make_grids(size, 2, seed=60)
params = I.make_mlp(73, seed=61, gain=1.5)
mip, fl, nc, crop = 2, 0, 2, 8
coord = np.array([[0, 0], [8, 16]])
N = nc * crop * crop
fp = [T(a) for a in grids]
pt = [T(p) for p in params]
m = L.make_mlp(pt)
g0t, g1t = fp[0], fp[1]
geom = L.make_geom(L.METHOD_2D, g0t, g1t, crop, nc, mip - 2, mip, 6, L.PE_TRIANGULAR)
h = L.handle(dev)
coord_t = T(coord, torch.int64)
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def run(target):
    g = [torch.zeros_like(p) for p in pt]
    gm = L.make_mlp_grad(g)
    d0, d1 = torch.zeros_like(g0t), torch.zeros_like(g1t)
    ls = torch.zeros(4, device=dev)
    o = torch.empty((N, 3), device=dev)
    tt = T(target)
    L.check(h, L.load_library().nic_train_step(h, C.byref(geom), L.ptr(g0t), L.ptr(g1t), L.ptr(coord_t), C.byref(m),
                                               L.ptr(tt), None, 0, 0, 0, 0, C.byref(gm), L.ptr(d0), L.ptr(d1),
                                               L.ptr(ls), L.ptr(o), L.PRECISIONS[prec], L.stream_ptr(dev)))
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in g], d0.cpu().numpy(), d1.cpu().numpy(), o.cpu().numpy()


_, _, _, out_tc = run(np.zeros((N, 3), np.float32))
_, out_or, _, _, _ = O.train_forward_backward(grids, params, coord, np.zeros((N, 3), np.float32), fl, mip, 1, None, size=crop)
for s in (0, 1, 7, 8, 31, 32, 33, 63, 64, 65, 100, 127):
    t_tc = out_tc.copy()
    t_tc[s] += 0.3
    t_or = out_or.astype(np.float32).copy()
    t_or[s] += 0.3
    g, d0, d1, _ = run(t_tc)
    _, _, grads, dg0, dg1 = O.train_forward_backward(grids, params, coord, t_or, fl, mip, 1, None, size=crop)
    print(f"s={s:3d}  W3 {rel(g[4], grads['W3']):.4f}  W2 {rel(g[2], grads['W2']):.4f}  b2 {rel(g[3], grads['b2']):.4f}  "
          f"W1 {rel(g[0], grads['W1']):.4f}  b1 {rel(g[1], grads['b1']):.4f}  dG0 {rel(d0, dg0):.4f}  dG1 {rel(d1, dg1):.4f}")
    if s in (0, 33):
        print("   b1 ratio:", np.round(g[1] / grads["b1"], 2))
