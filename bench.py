#!/usr/bin/env python
"""bench.py — headline benchmark of the decode hot path (BASELINE.json configs[1]):
full-frame decode of a 4096x4096 RGB texture on the f16 tensor-core path, one frame per GPU per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--prec f16|bf16|f32]

Prints ONE JSON line (see DESIGN.md "Measurement").  `value` = decoded Gtexel/s over all ranks with the
grids and decoder already resident in HBM; `e2e` = the same metric through the public API with host buffers
(pinned H2D of the compressed grids + decoder, unpack, decode, D2H of the 8-bit frame inside the timed
region).  `--impl reference` times the CPU port of the reference's algorithm (oracle/) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

SIZE = 4096                      # BASELINE.json configs[1]
FLOP_PER_TEXEL = 2 * (73 * 64 + 64 * 64 + 64 * 3)      # 17,920
METRIC, UNIT = "decoded Gtexel/s (4096x4096 RGB full-frame decode)", "Gtexel/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synthetic_model(seed=0):
    import inputs as I
    grids = I.make_grids(SIZE, 2, seed=seed, no_mip=True, quantized=True)      # [12,1025,1025], [12,513,513]
    params = I.make_mlp(73, seed=seed + 1, gain=2.0)
    return grids, params


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_decode_rate(rows, threads):
    """Times the oracle port (numpy restatement of the reference algorithm) on a slab of `rows` x 4096 texels."""
    from oracle import nic_oracle as O
    try:
        import torch
        torch.set_num_threads(threads)
    except Exception:
        pass
    grids, params = synthetic_model()
    table = O.create_pyramid_mip_levels(SIZE, SIZE // 4)
    chunk = 64
    t0 = time.perf_counter()
    done = 0
    for r0 in range(0, rows, chunk):
        blk = min(chunk, rows - r0)
        # a [blk, 4096] slab = blk single-row blocks would be slow in numpy; decode it as (blk x 4096) via two 1-D meshes
        xin = slab_input(O, grids, r0, blk)
        out = O.mlp_forward(xin, params)
        _ = O.quantize_to_bit(out, 8).astype(np.uint8)
        done += blk * SIZE
    dt = time.perf_counter() - t0
    return done / dt / 1e9, done, dt


def slab_input(O, grids, r0, rows):
    """Decoder input of the texel slab [r0, r0+rows) x [0, 4096) following oracle.decoder_input_one, for a
    non-square block (the reference only decodes squares; the arithmetic per texel is identical)."""
    F32 = np.float32
    ax_x = O._axis_vectors(r0, rows, 0.25)
    ax_y = O._axis_vectors(0, SIZE, 0.25)
    mesh = lambda a, b: [m.reshape(-1) for m in np.meshgrid(a, b, indexing="ij")]
    x0, y0 = mesh(ax_x[1], ax_y[1])
    x1, y1 = mesh(ax_x[3], ax_y[3])
    ux, uy = mesh(ax_x[2], ax_y[2])
    kx, ky = mesh(ax_x[4], ax_y[4])
    g0, g1 = grids[0], grids[1]
    rows_ = [g0[:, y0 + dy, x0 + dx] for dy, dx in O._CORNERS_2D]
    one = F32(1)
    wx = [one - kx, one - kx, kx, kx]
    wy = [one - ky, ky, one - ky, ky]
    g1c = [(g1[:, y1 + dy, x1 + dx] * wx[j]) * wy[j] for j, (dy, dx) in enumerate(O._CORNERS_2D)]
    rows_.append(((g1c[0] + g1c[1]) + g1c[2]) + g1c[3])
    rows_.append(O.triangular_positional_encoding(np.stack([ux, uy]), 6))
    rows_.append(np.zeros((1, x0.shape[0]), dtype=F32))
    return np.ascontiguousarray(np.concatenate(rows_, axis=0).T)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = args.cpu_rows
    for _ in range(min(args.warmup, 1)):
        cpu_decode_rate(16, threads)
    rates, tot_t = [], 0.0
    for _ in range(max(1, min(args.steps, 3))):
        r, done, dt = cpu_decode_rate(rows, threads)
        rates.append(r)
        tot_t += dt
    v = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * tot_t / len(rates), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "decode_4096x4096_rgb", "sample": f"{rows}x4096 texel slab of the frame per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{rows}x4096 texel slab per step, numpy port of the reference algorithm (oracle/nic_oracle.py)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import neural_image_compression_v2_b200 as nic
    from neural_image_compression_v2_b200 import image_compression as ic
    from neural_image_compression_v2_b200 import fp_def, var2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    var2.update(IMAGE_SIZE=SIZE)
    grids, params = synthetic_model(seed=rank)             # one frame per rank (frame-sharded decode, no collective)
    fp = [torch.tensor(g, device=dev) for g in grids]
    dec = ic.ColorDecoder(73, 64, 3).to(dev)
    with torch.no_grad():
        for p, v in zip(dec.parameters_list(), params):
            p.copy_(torch.tensor(v))
    out = torch.empty((SIZE, SIZE, 3), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    texels = SIZE * SIZE

    def step():
        ic.decode(fp, dec, 0, precision=args.prec, out_dtype=torch.uint8, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = nic.launch_count(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        s.record()
        step()
        e.record()
    barrier()
    launches = nic.launch_count(dev) - l0
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = world * texels / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers in, host buffer out, through the public API
    codes = [c.cpu().pin_memory() for c in fp_def.fp_savable(fp, 8)]
    host_params = [p.detach().cpu().pin_memory() for p in dec.parameters_list()]
    host_out = torch.empty((SIZE, SIZE, 3), dtype=torch.uint8).pin_memory()
    dcodes = [torch.empty_like(c, device=dev) for c in codes]
    h2d = sum(c.numel() for c in codes) + sum(p.numel() * 4 for p in host_params)
    d2h = host_out.numel()

    def e2e_step():
        for d, c in zip(dcodes, codes):
            d.copy_(c, non_blocking=True)
        with torch.no_grad():
            for p, hp in zip(dec.parameters_list(), host_params):
                p.copy_(hp, non_blocking=True)
        f = fp_def.fp_load(dcodes, 8)
        ic.decode(f, dec, 0, precision=args.prec, out_dtype=torch.uint8, out=out)
        host_out.copy_(out, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    k2 = max(3, min(args.steps, 10))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k2):
        e2e_step()
    e.record()
    barrier()
    ms2 = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * texels / (float(ms2.item()) / k2 * 1e-3) / 1e9

    if rank == 0:
        tf_peak, hbm_peak, src = peaks()
        achieved = FLOP_PER_TEXEL * texels / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.prec, "data": "synthetic",
            "config": {"workload": "decode_4096x4096_rgb", "frames_per_step": world, "grids": "[12,1025,1025]+[12,513,513] 8-bit",
                       "decoder": "73-64-64-3", "output": "uint8", "l2": "flushed between timed iterations (256 MB write)"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                         "traffic": None, "peak_source": f"{src} bf16_tflops (burst: kernel timed alone)",
                         "flop_per_texel": FLOP_PER_TEXEL},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            r, done, dt = cpu_decode_rate(args.cpu_rows, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{args.cpu_rows}x4096 texel slab ({dt:.1f} s), numpy port of the reference algorithm"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--prec", default="f16", choices=["f16", "bf16", "f32"])
    ap.add_argument("--cpu-rows", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
