#!/bin/bash
# One validation cycle on the GPU box, to be passed to gpurun as a single command (a call is charged acquire + push + run;
# this whole cycle is ~2 min):   gpurun --timeout 500 -- 'bash tools/gpu_cycle.sh <tag> [steps]'
# Writes gpurun_out/<tag>_{pytest.log,bench.json,bench.err,ref.json,launches.csv}; summarise here afterwards with
#   python tools/launch_summary.py gpurun_out/<tag>_launches.csv "<command>" > profiles/rNN_launches_<tag>.txt
# ncu --set full of one kernel (≈40 s):
#   ncu --set full --clock-control none --import-source on -k regex:<kernel> --launch-skip 2 -c 1 -o gpurun_out/<tag>_<k> -f \
#       python tools/prof_apply.py <nr> <E> <napply> <npcg> <k> <precond>      then: python tools/ncu_summary.py <rep>
tag=${1:-cycle}
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
( timeout 200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ); echo "bench exit $?"
( timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2>> gpurun_out/${tag}_bench.err ); echo "reference exit $?"
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --skip-cpu --skip-cfg4 --skip-cfg5 --skip-e2e --skip-cfg2 > /dev/null 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("value %.1f GDOF/s  %.4f ms/step  frac %.3f  e2e %.2f  clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
for k, v in d["extra"].items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("workload", "note", "l2")})
PY
