#!/bin/bash
# round 2, first GPU call: parity of the new PIP path, then timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_pip.py -x -q > gpurun_out/r2a_pytest_pip.log 2>&1; echo "pip tests rc=$?" >> gpurun_out/r2a_pytest_pip.log
tail -5 gpurun_out/r2a_pytest_pip.log
for cfg in "GPL_PIP_LEGACY=1" "GPL_PIP_RASTER_LOG2=6" "GPL_PIP_RASTER_LOG2=5" "GPL_PIP_RASTER_LOG2=7 GPL_PIP_RASTER_MCELLS=256" "GPL_PIP_RASTER_LOG2=6 GPL_L2_PIN=0" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_STREAM_CTAS_PER_SM=2" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_SLOTS_X100=600"; do
  env $cfg timeout 300 python tools/exp_pip2.py --tag "$cfg" >> gpurun_out/r2a_exp.jsonl 2>> gpurun_out/r2a_exp.err
done
env GPL_PIP_RASTER_LOG2=6 timeout 300 python tools/exp_pip2.py --outside --tag "outside" >> gpurun_out/r2a_exp.jsonl 2>> gpurun_out/r2a_exp.err
cat gpurun_out/r2a_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| walk',round(d['walk_cell_frac'],3),'| chk',d['checksum'],'| MB',round(d['index_MB'],1),'| def',d['deferred_per_query'])
"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_all.log 2>&1; tail -5 gpurun_out/r2a_pytest_all.log
