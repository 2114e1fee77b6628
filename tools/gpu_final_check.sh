#!/bin/bash
# Everything the round-end driver runs, plus the profile captures, in ONE gpurun call (run from the repo root on the GPU box):
#   ncu --set full capture of the hot kernel, ncu launch list of a short bench, pytest -m gpu, bench.py, smoke(), compute-sanitizer.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_final_check.sh <tag>'      (outputs under gpurun_out/)
tag=${1:-final}
NUM=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:raisr_frame_pipe -s 3 -c 1 -o gpurun_out/pipe_$tag -f python profiles/prof_one.py > gpurun_out/ncu_$tag.log 2>&1; tail -1 gpurun_out/ncu_$tag.log
RAISR_CUDA_NO_MEMOPS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/b_ncu_$tag.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_$tag.log 2>&1; tail -2 gpurun_out/t_$tag.log
timeout 500 python bench.py > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/b_$tag.json").read().strip().splitlines()[-1])
print("value",d["value"], "e2e",d["e2e"]["value"], "pageable",d["e2e_pageable"]["value"], "kernel_ms",d["roofline"]["kernel_ms"], "cpu", d["cpu_baseline"]["value"])
for k,v in d["configs"].items(): print(k, v.get("device_ms_per_frame"), v.get("e2e_frames_per_s"))
print(d["numerics_variants"])
PY
python -c "import __graft_entry__ as g; g.smoke()"
export RAISR_CUDA_NO_MEMOPS=1
(timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_probe.py; timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_probe.py 640 360 1) 2>&1 | grep -E "SUMMARY"
