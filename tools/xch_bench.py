"""Times nic_adam_step_exchange alone (one-shot and sliced) on synthetic flat gradient buffers of a given size:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/xch_bench.py [MB ...]
Prints, per buffer size and mode, the kernel time (CUDA events, max over ranks) and the NVLink bytes it moves per rank."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes_mb = [float(a) for a in sys.argv[1:] if not a.startswith("dbg=")] or [1.0, 15.8]
    dbg = sum(int(a[4:]) for a in sys.argv[1:] if a.startswith("dbg="))
    h, lib = L.handle(dev), L.load_library()
    token = 0
    for mb in sizes_mb:
        total = int(mb * 1e6 / 4) // 4 * 4
        buf = L.SymmetricBuffer(dev, 2 * total + 64)
        handles = [None] * world
        dist.all_gather_object(handles, buf.handle)
        bases = [buf.ptr if r == rank else buf.open_peer(handles[r]) for r in range(world)]
        flat0 = buf.tensor[:total]
        flat0.copy_(torch.randn(total, device=dev) * 1e-2)
        p, m, v = (torch.zeros(total - 4, device=dev) for _ in range(3))
        arr = (L.NicAdamTensor * 1)()
        t = arr[0]
        t.p, t.g, t.m, t.v, t.numel = p.data_ptr(), flat0.data_ptr(), m.data_ptr(), v.data_ptr(), total - 4
        t.lr, t.t, t.clamp, t.clamp_lo, t.clamp_hi = 0.01, 3, 1, -0.05, 0.05
        loss = torch.zeros(2, device=dev)
        nccl_buf = torch.randn(total, device=dev)
        if dbg:
            L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, dbg)
        for name, mode in (("one-shot", L.EXCHANGE_ONE_SHOT), ("sliced", L.EXCHANGE_SLICED), ("nccl all_reduce (no Adam)", None)):
            times = []
            for it in range(12):
                torch.cuda.synchronize()
                dist.barrier()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                if mode is None:
                    dist.all_reduce(nccl_buf)
                else:
                    token += 1
                    x = L.NicExchange()
                    x.world, x.rank, x.token, x.reserved = world, rank, token, mode
                    for r in range(world):
                        x.peer_flat[r], x.peer_flag[r] = bases[r], bases[r] + 8 * total
                    x.zero_buf, x.zero_numel = buf.ptr + 4 * total, total
                    L.check(h, lib.nic_adam_step_exchange(h, arr, 1, 0.9, 0.999, 1e-8, 1.0, C.byref(x), L.ptr(flat0[total - 4:]),
                                                          L.ptr(loss), 0.5, L.stream_ptr(dev)))
                e.record()
                torch.cuda.synchronize()
                times.append(s.elapsed_time(e))
            tt = torch.tensor([float(np.median(times[2:]))], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if mode is not None:
                assert not L.exchange_status(dev)
            if rank == 0:
                moved = {"one-shot": (world - 1) * total * 4, "sliced": 2 * (world - 1) * total * 4 / world}.get(name, 0)
                print(f"XCH world {world} {mb:5.1f} MB {name:26s} {tt.item() * 1e3:8.1f} us  NVLink bytes/rank {moved / 1e6:6.1f} MB"
                      + (f"  ({moved / tt.item() / 1e6:.0f} GB/s incl. handshakes and the dense Adam)" if moved else ""), flush=True)
        if dbg:
            L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, 0)
        dist.barrier()
        del flat0
        buf.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
