python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_v7.txt
python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench768_v7.json 2> gpurun_out/bench768_v7.err
SKIT_UNPACK_TILED=0 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench768_v7_notiled.json 2>> gpurun_out/bench768_v7.err
python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench768_v7b.json 2>> gpurun_out/bench768_v7.err
cat gpurun_out/tests_v7.txt; python - <<'PY'
import json
for f in ("bench768_v7","bench768_v7_notiled","bench768_v7b"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
