#!/bin/bash
# vocab-parallel logits over NCCL on N GPUs (BASELINE configs[4]): tests (N = 2) + bench lines.  usage: r2_vocab_parallel.sh N tag
N=${1:-2}; tag=${2:-r2j}
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -25 > gpurun_out/${tag}_tests.log; fi
port=29600
for s in greedy topk nucleus; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --mode vocab-parallel --sampling $s --steps 50 --warmup 5 > gpurun_out/${tag}_vp${N}_${s}.json 2> gpurun_out/${tag}_vp${N}_${s}.err
done
python - <<PY
import json
for s in ("greedy", "topk", "nucleus"):
    f = "gpurun_out/${tag}_vp${N}_%s.json" % s
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(s, "N", d["n_gpus"], round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "allgather", d["allgather"], "same tokens", d["tokens_identical_on_all_ranks"])
    except Exception as ex:
        print(f, "failed:", ex)
PY
