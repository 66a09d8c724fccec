#!/bin/bash
# round 2, second GPU call: new raster build, register variants, bench, ncu of the shipped kernel, full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2b
timeout 900 python -m pytest tests/test_gpu_pip.py -x -q > gpurun_out/${T}_pytest_pip.log 2>&1; echo "pip tests rc=$?" >> gpurun_out/${T}_pytest_pip.log
tail -4 gpurun_out/${T}_pytest_pip.log
rm -f gpurun_out/${T}_exp.jsonl
for cfg in "GPL_PIP_RASTER_LOG2=6" "GPL_PIP_RASTER_LOG2=5" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_STREAM_MINB=4" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_FUSE_HIST=1" "GPL_PIP_RASTER_LOG2=6 GPL_PIP_SLOTS_X100=200"; do
  env $cfg timeout 300 python tools/exp_pip2.py --tag "$cfg" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
done
env GPL_PIP_RASTER_LOG2=7 timeout 300 python tools/exp_pip2.py --polys 1000 --grid 32 --cell 31.25 --points 125000000 --tag "c4 log2=7" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
env GPL_PIP_RASTER_LOG2=6 timeout 300 python tools/exp_pip2.py --polys 1000 --grid 32 --cell 31.25 --points 125000000 --tag "c4 log2=6" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
cat gpurun_out/${T}_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| walk',round(d['walk_cell_frac'],3),'| chk',d['checksum'],'| MB',round(d['index_MB'],1),'| def',d['deferred_per_query'])
"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 2500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
# launch list of the bench command (shares of the step) and one full capture of the streaming kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_stream -s 3 -c 1 -o gpurun_out/${T}_pip_stream python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_pip_stream.ncu-rep
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_all.log 2>&1; tail -6 gpurun_out/${T}_pytest_all.log
