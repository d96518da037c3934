#!/bin/bash
# round 2: N-GPU SILK bench (BASELINE configs[2] per GPU, weak scaling, one process per GPU, no data-path collective)
# usage: tools/gpu/r2s.sh N [steps] [warmup]
N=$1; K=${2:-100}; W=${3:-5}
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --silk --gpus $N --steps $K --warmup $W --no-cpu-baseline > $O/r2s_silk_${N}gpu.json 2> $O/r2s_silk_${N}gpu.err
tail -3 $O/r2s_silk_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_silk_${N}gpu.json"))
print("SILK N=${N}", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/r2s_celt_${N}gpu.json 2> $O/r2s_celt_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_celt_${N}gpu.json"))
print("CELT N=${N}", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
PY
