"""A/B sweep of NIC_OPT_GELU_POLY on the fast 2-D decode kernel: kernel time at 4096^2 and parity of the 8-bit frame
against the reference-exact fp32 path at 1024^2.   python tools/gelu_sweep.py [npoly,npoly,...]
(values other than 0 and the default need a library built with -DNIC_GELU_SWEEP)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

dev = torch.device("cuda:0")
sweep = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 3]
for a in sys.argv[2:]:
    if a.startswith("dbg="):
        L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, int(a[4:]))


def model(size, seed, gain):
    var2.update(IMAGE_SIZE=size)
    fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=seed, no_mip=True, quantized=True)]
    dec = ic.ColorDecoder(73, 64, 3).to(dev)
    with torch.no_grad():
        for p, v in zip(dec.parameters_list(), I.make_mlp(73, seed=seed + 1, gain=gain)):
            p.copy_(torch.tensor(v))
    return fp, dec


def psnr(a, b):
    mse = ((a.float() - b.float()) ** 2).mean().item()
    return 10 * torch.log10(torch.tensor(256.0 ** 2 / max(mse, 1e-12))).item()


small = [(model(1024, 10, 2.0)), (model(1024, 20, 4.0))]
refs = []
for fp, dec in small:
    var2.update(IMAGE_SIZE=1024)
    r32 = ic.decode(fp, dec, 0, precision="f32")
    refs.append((r32, torch.floor(r32 * 255 + 0.5).to(torch.uint8)))
big = model(4096, 0, 2.0)
out = torch.empty((4096, 4096, 3), dtype=torch.uint8, device=dev)
for prec in ("f16", "bf16"):
    for npoly in sweep:
        L.set_option(dev, L.OPT_GELU_POLY, npoly)
        stats = []
        for (fp, dec), (r32, r8) in zip(small, refs):
            var2.update(IMAGE_SIZE=1024)
            u8 = ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8)
            d = (u8.int() - r8.int()).abs()
            tex = d.amax(dim=-1)
            # PSNR of each 8-bit frame against a pseudo ground truth (the fp32 float output): the delta is what matters
            gt = (r32 * 255).clamp(0, 255)
            stats.append((float((tex == 0).float().mean()), float((tex <= 1).float().mean()), int(d.max()),
                          psnr(u8, gt) - psnr(r8, gt)))
        var2.update(IMAGE_SIZE=4096)
        fp, dec = big
        for _ in range(2):
            ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8, out=out)
        torch.cuda.synchronize()
        L.set_option(dev, L.OPT_TIME_KERNELS, 1)
        for _ in range(8):
            ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8, out=out)
        torch.cuda.synchronize()
        ms, n = L.kernel_time_ms(dev)
        L.set_option(dev, L.OPT_TIME_KERNELS, 0)
        print(f"{prec} npoly={npoly}: {ms / n:.4f} ms  {4096 * 4096 / (ms / n * 1e-3) / 1e9:.2f} Gtexel/s | " +
              " | ".join(f"identical {a:.4f} within1 {b:.5f} max {c} dPSNR {e:+.4f}" for a, b, c, e in stats), flush=True)
L.set_option(dev, L.OPT_GELU_POLY, -1)
