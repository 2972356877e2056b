"""Small driver for profiling the fused training step: python tools/run_train.py [f16|bf16|f32] [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
size, nc, crop = 512, 8, 256
var2.update(IMAGE_SIZE=size)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=3, no_mip=True)]
dec = ic.ColorDecoder(73, 64, 3).to(dev)
with torch.no_grad():
    for p, v in zip(dec.parameters_list(), I.make_mlp(73, seed=4)):
        p.copy_(torch.tensor(v))
tr = ic.FusedTrainer(fp, dec, num_epochs=100000, fp_bits=8, seed=1, precision=prec)
img = torch.tensor(I.make_image(size, 2, seed=5), device=dev)
g = torch.Generator().manual_seed(100)
coord = torch.randint(0, size - crop + 1, (nc, 2), generator=g).to(dev)
tg = ic.sample_crops(img, coord, crop)
noise_arg = False if "nonoise" in sys.argv else None
if "frozen" in sys.argv:
    tr.frozen = True
dbg_extra = sum(int(a[4:]) for a in sys.argv if a.startswith("dbg="))      # e.g. dbg=256: shuffle scatter instead of the MMA scatter
if dbg_extra:
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, dbg_extra)
if "static" in sys.argv:            # static round-robin tile order instead of the atomic counter
    L.set_option(dev, L.OPT_STATIC_TILES, 1)
for _ in range(3):
    tr.step(coord, tg, 0, noise=noise_arg)
torch.cuda.synchronize()
if "notimer" not in sys.argv:       # the library's events around the training kernel rule out its programmatic launch
    L.set_option(dev, L.OPT_TIME_KERNELS, 1)
t0 = time.perf_counter()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(steps):
    loss = tr.step(coord, tg, 0, noise=noise_arg)
e.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
ms, n = L.kernel_time_ms(dev)
print(f"{prec}: step {s.elapsed_time(e) / steps:.4f} ms (GPU events), host issue {1e3 * (t1 - t0) / steps:.4f} ms/step, "
      f"train kernel {ms / max(n, 1):.4f} ms, {nc * crop * crop / (s.elapsed_time(e) / steps * 1e-3) / 1e6:.0f} Msamples/s, loss {float(loss):.6f}")

if "phases" in sys.argv:
    # per-tile phase profile of the tensor-core training kernel as seen by thread 0 of every CTA (nic_debug_counters)
    extra = sum(int(a[4:]) for a in sys.argv if a.startswith("dbg="))
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, 8 | extra)
    tr.step(coord, tg, 0, noise=noise_arg)
    L.debug_counters(dev)
    for _ in range(5):
        tr.step(coord, tg, 0, noise=noise_arg)
    c = L.debug_counters(dev)
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, 0)
    names = ["gather", "mma L1", "gelu 1", "mma L2", "gelu 2", "mma L3", "loss/dz3", "mma dH2+D3", "dz2", "mma dH1+D2", "dz1",
             "mma dX+D1", "scatter", "flush (x tiles/CTA)", "CTA lifetime (x tiles/CTA)"]
    tiles = max(c[15], 1)
    tot = sum(c[:13])
    print(f"phase profile: {tiles} tiles, {tot / tiles:.0f} cycles per tile per CTA")
    for i, nme in enumerate(names):
        print(f"  {nme:12s} {c[i] / tiles:8.0f} cycles  {100.0 * c[i] / tot:5.1f} %")

if "timeline" in sys.argv:
    # nanosecond stamps (globaltimer) of one launch: CTA entry / first tile / flush start / exit, min and max over the CTAs
    L.set_option(dev, L.OPT_TIME_KERNELS, 1)
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, 8 | 512 | dbg_extra)
    tr.step(coord, tg, 0, noise=noise_arg)
    L.debug_counters(dev)
    for rep in range(3):
        tr.step(coord, tg, 0, noise=noise_arg)
        c = [x & 0xFFFFFFFFFFFFFFFF for x in L.debug_counters(dev)]
        inv = lambda v: (~v) & 0xFFFFFFFFFFFFFFFF
        t0 = inv(c[0])
        names = ["entry min", "entry max", "loop min", "loop max", "flush min", "flush max", "exit min", "exit max"]
        vals = [inv(c[0]), c[1], inv(c[2]), c[3], inv(c[4]), c[5], inv(c[6]), c[7]]
        print("timeline (us from the first CTA entry): " + ", ".join(f"{n} {(v - t0) / 1e3:.1f}" for n, v in zip(names, vals)))
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, 0)
