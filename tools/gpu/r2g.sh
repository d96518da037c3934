#!/bin/bash
# round 2: N-GPU bench (weak scaling, one process per GPU, NCCL only for the barrier / MAX of the elapsed time)
# usage: tools/gpu/r2g.sh N [steps] [warmup]
N=$1; K=${2:-100}; W=${3:-5}
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $K --warmup $W --no-cpu-baseline > $O/r2g_bench_${N}gpu.json 2> $O/r2g_bench_${N}gpu.err
tail -3 $O/r2g_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_${N}gpu.json"))
print("N=${N}", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "pcie_peak", round(d["e2e"]["pcie_peak_gbs"],1), "pcie_frac", round(d["e2e"]["pcie_frac"],3), "i16", round(d["e2e_i16"]["value"]), "cpus", d["detail"]["rank_cpus"])
PY
nvidia-smi topo -m > $O/r2g_topo_${N}gpu.txt 2>&1; head -14 $O/r2g_topo_${N}gpu.txt; nproc; lscpu | grep -E "NUMA|Socket|Model name" | head
