python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_v6.txt
python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench768_v6.json 2> gpurun_out/bench768_v6.err
SKIT_WGRAD_CO1=0 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench768_v6_noco1.json 2>> gpurun_out/bench768_v6.err
python bench.py --no-cpu-baseline --no-extra --size 512 --no-nce > gpurun_out/bench512_v6.json 2>> gpurun_out/bench768_v6.err
cat gpurun_out/tests_v6.txt; python - <<'PY'
import json
for f in ("bench768_v6","bench768_v6_noco1","bench512_v6"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"])
PY
