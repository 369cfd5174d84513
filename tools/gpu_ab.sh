#!/bin/bash
# A/B on one GPU box: parity tests of the TRX path, then short bench runs under different switches.
# usage: tools/gpu_ab.sh <tag> [VAR=VALUE ...]   (each extra argument is one more bench variant)
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trx.py tests/test_gpu_trx_attn.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --global-episodes 1024 --no-cpu-baseline --no-extras \
      > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms/micro", round(d["config"]["ms_per_micro_batch"], 3), "gemm_ms", round(d["roofline"]["kernel_ms_per_micro_batch"], 3),
          "frac", round(d["roofline"]["frac"], 4), "tuple_ms", round(d["roofline_tuple"]["kernel_ms_per_micro_batch"], 3),
          "tuple_frac", round(d["roofline_tuple"]["frac"], 3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name}", "FAILED", e)
PY
}
run base LMKD_NOP=1
for v in "$@"; do
  run "$(echo $v | tr '=,' '__')" $(echo $v | tr ',' ' ')
done
