"""Data-parallel training check for >= 2 GPUs (launched by tests/test_gpu_parity.py::test_dp_exchange_* or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dp_exchange_check.py
Trains the same model with the exchange step as an NCCL all-reduce and fused into the optimiser over NVLink peer memory
(nic_adam_step_exchange), from the same initial state, crops and noise, and checks that
  * the replicas stay bit-identical across ranks in both modes,
  * the two modes agree to float rounding (the sums are formed in a different order),
  * no exchange timed out,
  * N ranks x b crops == ONE rank x N b crops (fp32 path, no noise): the data-parallel trajectory equals the single-device
    trajectory on the concatenated batch to float rounding,
  * a lost peer is FATAL, not silent: when one rank skips a step the others time out, apply no update, and their next
    step raises NIC_ERR_EXCHANGE.
Prints one line `DP_EXCHANGE_OK ...` on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    size, nc = 512, 4
    var2.update(IMAGE_SIZE=size, TF_NO_MIP=False, MAX_MIP_LEVEL=8)
    mips = [torch.tensor(m, device=dev) for m in I.box_mips(I.make_image(size, 2, seed=5), 8)]
    rng = np.random.default_rng(100 + rank)                 # every rank trains on its own crops
    lods = [0, 2, 0, 1, 2, 0, 0, 1, 3, 0, 1, 0, 2, 0]      # same on every rank; several pyramid levels, both parities
    batches = []
    for lod in lods:
        crop, dsize = 2 ** (8 - lod), size >> lod
        coord = torch.tensor(rng.integers(0, dsize - crop + 1, (nc, 2)))
        batches.append((coord, ic.sample_crops(mips[lod], coord, crop), lod))
    results = {}
    for mode in ("nccl", "peer", "sliced"):
        fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=3)]
        dec = ic.ColorDecoder(73, 64, 3).to(dev)
        with torch.no_grad():
            for p, v in zip(dec.parameters_list(), I.make_mlp(73, seed=4)):
                p.copy_(torch.tensor(v))
        tr = ic.FusedTrainer(fp, dec, num_epochs=1000, fp_bits=8, seed=1, precision=sys.argv[1] if len(sys.argv) > 1 else "f16",
                             exchange=mode)
        losses = []
        for coord, tg, lod in batches:
            losses.append(tr.step(coord, tg, lod))
        torch.cuda.synchronize()
        assert not L.exchange_status(dev), "an exchange timed out"
        if mode != "nccl":
            assert tr.exchange_in_use() == mode, f"{mode} exchange was not used: {tr.exchange_in_use()}"
        state = torch.cat([g.reshape(-1) for g in fp] + [p.detach().reshape(-1) for p in dec.parameters_list()])
        gathered = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(gathered, state)
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), f"{mode}: replicas of rank 0 and {r} differ"
        results[mode] = (state.cpu().numpy(), np.array([float(x) for x in losses]))
    a, b = results["nccl"], results["peer"]
    rel = float(np.linalg.norm(a[0] - b[0]) / np.linalg.norm(a[0]))
    assert rel < 2e-3, f"nccl and peer exchange disagree: rel {rel}"
    assert np.allclose(a[1], b[1], rtol=2e-2, atol=1e-5), (a[1], b[1])
    # the sliced exchange against nccl to the same tolerance (two training runs differ in the last bits anyway: the order
    # of the scatter's float atomics), and against the one-shot exchange BIT FOR BIT on the same gradient buffers
    c = results["sliced"]
    rel_s = float(np.linalg.norm(a[0] - c[0]) / np.linalg.norm(a[0]))
    assert rel_s < 2e-3, f"nccl and sliced exchange disagree: rel {rel_s}"
    assert np.allclose(a[1], c[1], rtol=2e-2, atol=1e-5), (a[1], c[1])
    exchange_kernel_modes_identical(rank, world, dev)
    rel1 = dp_equals_single_device(rank, world, dev, mips)
    fatal_timeout(rank, world, dev, mips, mode="peer")
    fatal_timeout(rank, world, dev, mips, mode="sliced")
    dist.barrier()
    if rank == 0:
        print(f"DP_EXCHANGE_OK world {world} rel_l2(nccl, peer) {rel:.2e} final loss {b[1][-1]:.5f} "
              f"rel_l2(nccl, sliced) {rel_s:.2e} rel_l2(dp, single device) {rel1:.2e} sliced == one-shot kernel bit for bit; lost-peer timeout fatal: yes (both modes)")
    dist.destroy_process_group()


def exchange_kernel_modes_identical(rank, world, dev):
    """nic_adam_step_exchange on the SAME per-rank gradient buffers, one-shot and sliced: rank r reduces slice r in rank
    order and writes it to every buffer, so parameters, Adam state and loss must come out bit-identical to the one-shot sum,
    and every rank's buffer must end up holding the same reduced gradient."""
    import ctypes as C
    sizes = [12 * 129 * 129, 12 * 65 * 65, 73 * 64, 64, 4]          # two grids, W1, b1, the loss slot
    offs, total = [], 0
    for sz in sizes:
        offs.append(total)
        total += (sz + 3) // 4 * 4
    buf = L.SymmetricBuffer(dev, 2 * total + 64)
    handles = [None] * world
    dist.all_gather_object(handles, buf.handle)
    bases = [buf.ptr if r == rank else buf.open_peer(handles[r]) for r in range(world)]
    flat0, flat1 = buf.tensor[:total], buf.tensor[total:2 * total]
    grad = torch.tensor(np.random.default_rng(50 + rank).standard_normal(total).astype(np.float32) * 1e-2, device=dev)
    flat0.copy_(grad)
    torch.cuda.synchronize()
    dist.barrier()
    h, lib = L.handle(dev), L.load_library()
    out = {}
    for token, mode in ((1, L.EXCHANGE_ONE_SHOT), (2, L.EXCHANGE_SLICED)):
        nt = len(sizes) - 1
        params = [torch.tensor(np.random.default_rng(60 + k).standard_normal(sizes[k]).astype(np.float32) * 0.1, device=dev)
                  for k in range(nt)]
        ms = [torch.full_like(p, 0.01) for p in params]
        vs = [torch.full_like(p, 0.001) for p in params]
        arr = (L.NicAdamTensor * nt)()
        for k in range(nt):
            t = arr[k]
            t.p, t.g = params[k].data_ptr(), flat0[offs[k]:].data_ptr()
            t.m, t.v, t.numel = ms[k].data_ptr(), vs[k].data_ptr(), sizes[k]
            t.lr, t.t, t.clamp, t.clamp_lo, t.clamp_hi = 0.01, 3, int(k < 2), -0.05, 0.05
        loss = torch.zeros(1, device=dev)
        x = L.NicExchange()
        x.world, x.rank, x.token, x.reserved = world, rank, token, mode
        for r in range(world):
            x.peer_flat[r], x.peer_flag[r] = bases[r], bases[r] + 8 * total
        x.zero_buf, x.zero_numel = buf.ptr + 4 * total, total
        flat1.fill_(7.0)
        L.check(h, lib.nic_adam_step_exchange(h, arr, nt, 0.9, 0.999, 1e-8, 1.0, C.byref(x), L.ptr(flat0[offs[nt]:]), L.ptr(loss),
                                              0.5, L.stream_ptr(dev)))
        torch.cuda.synchronize()
        assert not L.exchange_status(dev) and float(flat1.abs().sum()) == 0.0
        out[mode] = [t.cpu().numpy() for t in params + ms + vs] + [loss.cpu().numpy()]
        if mode == L.EXCHANGE_ONE_SHOT:
            assert torch.equal(flat0, grad)                      # one-shot leaves the gradient buffers alone
    for u, v in zip(out[L.EXCHANGE_ONE_SHOT], out[L.EXCHANGE_SLICED]):
        assert np.array_equal(u, v), "one-shot and sliced exchange kernels differ"
    dist.barrier()
    reduced = [torch.empty_like(flat0) for _ in range(world)]
    dist.all_gather(reduced, flat0.clone())
    for r in range(1, world):
        assert torch.equal(reduced[0], reduced[r]), "sliced exchange: the reduced gradient differs between ranks"
    del flat0, flat1
    dist.barrier()
    buf.close()


def fresh_model(dev, size):
    fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=3)]
    dec = ic.ColorDecoder(73, 64, 3).to(dev)
    with torch.no_grad():
        for p, v in zip(dec.parameters_list(), I.make_mlp(73, seed=4)):
            p.copy_(torch.tensor(v))
    return fp, dec


def dp_equals_single_device(rank, world, dev, mips, size=512, nc=2, steps=6):
    """Every rank: `steps` fp32 steps, data parallel on its own crops AND alone on the crops of ALL ranks."""
    rng = np.random.default_rng(7)
    lods = [0, 1, 0, 2, 1, 0][:steps]
    out = {}
    for mode in ("dp", "single"):
        fp, dec = fresh_model(dev, size)
        tr = ic.FusedTrainer(fp, dec, num_epochs=1000, fp_bits=8, seed=1, precision="f32", exchange="peer",
                             data_parallel=mode == "dp")
        r2 = np.random.default_rng(8)
        for lod in lods:
            crop, dsize = 2 ** (8 - lod), size >> lod
            allc = torch.tensor(r2.integers(0, dsize - crop + 1, (world * nc, 2)))       # the same draw on every rank
            coord = allc[rank * nc:(rank + 1) * nc] if mode == "dp" else allc
            tr.step(coord, ic.sample_crops(mips[lod], coord, crop), lod, noise=False)
        torch.cuda.synchronize()
        out[mode] = torch.cat([g.reshape(-1) for g in fp] + [p.detach().reshape(-1) for p in dec.parameters_list()]).cpu().numpy()
        if mode == "dp":
            dist.barrier()
            tr.close()
    rel = float(np.linalg.norm(out["dp"] - out["single"]) / np.linalg.norm(out["single"]))
    assert rel < 1e-5, f"data parallel and single-device trajectories differ: rel {rel}"
    return rel


def fatal_timeout(rank, world, dev, mips, size=512, nc=2, mode="peer"):
    """The last rank skips one step.  Every other rank must time out (300 ms here), leave its parameters untouched, and get
    NIC_ERR_EXCHANGE from its next step."""
    fp, dec = fresh_model(dev, size)
    tr = ic.FusedTrainer(fp, dec, num_epochs=1000, fp_bits=8, seed=1, precision="f16", exchange=mode, exchange_timeout_ms=300)
    coord = torch.tensor(np.random.default_rng(9 + rank).integers(0, size - 256 + 1, (nc, 2)))
    tg = ic.sample_crops(mips[0], coord, 256)
    tr.step(coord, tg, 0)                                   # a good step (maps the peer buffers)
    torch.cuda.synchronize()
    dist.barrier()
    before = torch.cat([g.reshape(-1) for g in fp] + [p.detach().reshape(-1) for p in dec.parameters_list()]).clone()
    if rank != world - 1:
        tr.step(coord, tg, 0)                               # the peer never arrives
        torch.cuda.synchronize()
        after = torch.cat([g.reshape(-1) for g in fp] + [p.detach().reshape(-1) for p in dec.parameters_list()])
        assert torch.equal(before, after), "a timed-out exchange must not update the parameters"
        try:
            tr.step(coord, tg, 0)
            raise AssertionError("the step after a timed-out exchange must raise")
        except L.NicError as e:
            assert e.status == L.ERR_EXCHANGE, e
        assert L.exchange_status(dev), "nic_exchange_status must report the timeout"
        assert not L.exchange_status(dev), "... and clear it"
    dist.barrier()
    tr._close_buffers()
    L.set_option(dev, L.OPT_EXCHANGE_TIMEOUT_MS, 0)


if __name__ == "__main__":
    main()
