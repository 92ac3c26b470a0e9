#!/bin/bash
# End-of-round verification on one B200 (gpurun): GPU tests, both bench arms, launch list, kernel counters, sanitizer,
# timeline.  Most important first: the call may be cut by its time limit.  usage: tools/final_check.sh <prefix>
p=gpurun_out/${1:-final}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q --durations=5 > ${p}_pytest.log 2>&1; tail -3 ${p}_pytest.log
timeout 120 python bench.py --impl reference > ${p}_bench_ref.json 2> ${p}_bench_ref.err; cat ${p}_bench_ref.json | cut -c1-300
timeout 200 python bench.py > ${p}_bench.json 2> ${p}_bench.err; cat ${p}_bench.json | cut -c1-400
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${p}_ncu_launch_list.csv python bench.py --steps 2 --warmup 1 > ${p}_ncu_bench.log 2>&1
timeout 90 bash tools/ncu_light.sh ${p}_picture_light.csv picture python tools/prof_run.py 3000 2 > ${p}_light1.log 2>&1
timeout 90 bash tools/ncu_light.sh ${p}_entropy_light.csv entropy python tools/prof_run.py 3000 2 > ${p}_light2.log 2>&1
timeout 90 compute-sanitizer --tool memcheck python tools/gpu_sanity.py > ${p}_memcheck.log 2>&1; tail -2 ${p}_memcheck.log
timeout 90 compute-sanitizer --tool racecheck python tools/gpu_sanity.py > ${p}_racecheck.log 2>&1; tail -2 ${p}_racecheck.log
timeout 60 python tools/e2e_timeline.py 3000 3 > ${p}_e2e_timeline.jsonl 2>/dev/null
HWB_TRACE_BATCHES=1 timeout 60 python tools/e2e_run.py 3000 > ${p}_e2e_trace.txt 2>&1
