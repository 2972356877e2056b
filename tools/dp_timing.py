"""Times the data-parallel training step with both exchange modes (torchrun, >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_timing.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
size, nc, crop, steps = 512, 8, 256, 100
var2.update(IMAGE_SIZE=size)
img = torch.tensor(I.make_image(size, 2, seed=5), device=dev)
g = torch.Generator().manual_seed(100 + rank)
coord = torch.randint(0, size - crop + 1, (nc, 2), generator=g).to(dev)
tg = ic.sample_crops(img, coord, crop)
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, int(os.environ.get("NIC_DBG", "0")))
for mode in ("nccl", "peer", "nccl", "peer"):
    fp = [torch.tensor(a, device=dev) for a in I.make_grids(size, 2, seed=3, no_mip=True)]
    dec = ic.ColorDecoder(73, 64, 3).to(dev)
    tr = ic.FusedTrainer(fp, dec, num_epochs=100000, fp_bits=8, seed=1, precision="f16", exchange=mode)
    for _ in range(5):
        tr.step(coord, tg, 0)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        tr.step(coord, tg, 0)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} exchange {mode}: {float(ms):.4f} ms/step")
dist.destroy_process_group()
