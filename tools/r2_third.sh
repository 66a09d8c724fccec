#!/bin/bash
# round 2, third GPU call: join / contains / distance tests, build restructure, smem candidate table, other bench arms at N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2c
timeout 1200 python -m pytest tests/test_gpu_pip.py tests/test_gpu_join.py tests/test_gpu_hull.py tests/test_gpu_predicates.py tests/test_gpu_ops.py tests/test_gpu_geodesic.py tests/test_gpu_formats.py -x -q > gpurun_out/${T}_pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_a.log
tail -15 gpurun_out/${T}_pytest_a.log
rm -f gpurun_out/${T}_exp.jsonl
for cfg in "GPL_PIP_RASTER_LOG2=6" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_CAND_SMEM=0" "GPL_PIP_RASTER_LOG2=5" "GPL_PIP_RASTER_LOG2=6 GPL_L2_PIN=0"; do
  env $cfg timeout 300 python tools/exp_pip2.py --tag "$cfg" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
done
env GPL_PIP_RASTER_LOG2=7 timeout 300 python tools/exp_pip2.py --polys 1000 --grid 32 --cell 31.25 --points 125000000 --tag "c4 log2=7" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
cat gpurun_out/${T}_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| walk',round(d['walk_cell_frac'],3),'| chk',d['checksum'],'| MB',round(d['index_MB'],1),'| def',d['deferred_per_query'])
"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; tail -c 1800 gpurun_out/${T}_bench_c2.json; tail -3 gpurun_out/${T}_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ncu_bench.log 2>&1
for w in c4 c3 c5; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "== $w rc=$?"; tail -c 1500 gpurun_out/${T}_bench_$w.json; tail -3 gpurun_out/${T}_bench_$w.err
done
