set -x
mkdir -p gpurun_out
python bench.py --steps 600 --warmup 20 --no-config2 --no-cpu > gpurun_out/zc_base.json 2> gpurun_out/zc_base.err
for c in 37 74 148 296; do
TRTLAB_ZERO_COPY_INPUT=1 TRTLAB_ZERO_COPY_CTAS=$c python bench.py --steps 600 --warmup 20 --no-config2 --no-cpu > gpurun_out/zc_$c.json 2> gpurun_out/zc_$c.err
done
python bench.py --steps 600 --warmup 20 --no-config2 --no-cpu > gpurun_out/zc_base2.json 2> gpurun_out/zc_base2.err
TRTLAB_ZERO_COPY_INPUT=1 timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/zc_tests.log 2>&1
tail -3 gpurun_out/zc_tests.log
for f in gpurun_out/zc_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d["e2e"]; print(" value",round(d["value"]), "e2e",round(e["value"]), "bracketed", round(e.get("bracketed",0)), "p50/p99", e.get("latency_ms_p50"), e.get("latency_ms_p99"))
except Exception as ex: print("ERR",ex)
PY
done
