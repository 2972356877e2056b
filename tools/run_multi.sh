#!/bin/bash
# Multi-GPU evidence run: tools/run_multi.sh N TAG  (writes gpurun_out/TAG_*.{log,json})
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1; nproc >> gpurun_out/${TAG}_topo.txt; numactl -H >> gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 $TR --master-port 29533 tests/dp_exchange_check.py > gpurun_out/${TAG}_dp_exchange_check.log 2>&1; grep -h "DP_EXCHANGE_OK\|Error\|assert" gpurun_out/${TAG}_dp_exchange_check.log | tail -5
timeout 900 $TR --master-port 29541 bench.py --gpus $N --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
for w in decode_4096_strong train_2048_mips lut_65 video_1080p120; do
  timeout 600 $TR --master-port 29547 bench.py --gpus $N --workload $w > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err || tail -5 gpurun_out/${TAG}_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "UNREADABLE", e); continue
    keep={k:d.get(k) for k in ("value","unit","ms_per_step","bands_equal_full_frame","replicas_identical","exchange","points_equal_dense_decode")}
    if "e2e" in d: keep["e2e"]=d["e2e"]["value"]; keep["floor"]=d["e2e"]["pcie_floor"]["gtexel_s"]; keep["floor_agg_gbs"]=d["e2e"]["pcie_floor"]["aggregate_gbs"]
    if "train" in d: keep["train"]={k:d["train"].get(k) for k in ("value","ms_per_step","ms_per_step_presampled","replicas_identical","exchange_timed_out")}
    print(f.split("/")[-1], keep)
PY
