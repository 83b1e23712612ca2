#!/bin/bash
# One GPU session: tests, smoke, bench, tile sweep.  Usage: gpurun --timeout N -- 'bash tools/gpu_round.sh [stage...]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build exit $?"; tail -2 gpurun_out/build.log
STAGES="${@:-tests smoke bench sweep}"
for st in $STAGES; do
case $st in
tests)
  timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
  grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40 ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d["e2e"]["value"], "roofline", d["roofline"], "\nbackbone", d["roofline_backbone"])
    print("cpu", d.get("cpu_baseline"))
    for l in d["layers"]:
        print(l["kernel"], round(l["ms"], 4), "ms", round(l["frac"], 3), "of HBM", round(l["tflops"], 1), "TF")
except Exception as e:
    print("bench parse failed", e)
PY
  tail -5 gpurun_out/bench.err ;;
sweep)
  timeout 900 python tools/tile_sweep.py 96 2048 > gpurun_out/tile_sweep_96.log 2>&1; echo "sweep exit $?"; tail -20 gpurun_out/tile_sweep_96.log ;;
sweep128)
  timeout 900 python tools/tile_sweep.py 128 1024 > gpurun_out/tile_sweep_128.log 2>&1; echo "sweep128 exit $?"; tail -20 gpurun_out/tile_sweep_128.log ;;
refbench)
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "refbench exit $?"; cat gpurun_out/bench_ref.json ;;
ncu)
  python tools/profile_target.py 96 2048 > gpurun_out/profile_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py 96 2048 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
  python tools/profile_target.py 96 2048 > gpurun_out/profile_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"stem_kernel|blaze_block" -c 17 -o gpurun_out/prof_backbone python tools/profile_target.py 96 2048 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
  ls -la gpurun_out/*.ncu-rep ;;
esac
done
