#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2j}
export GPL_HULL_FUSED=0
timeout 900 python -m pytest tests/test_gpu_hull.py tests/test_gpu_pip.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
for cfg in "GPL_HULL_MINB=5"; do
  echo "== $cfg"; env $cfg timeout 600 python bench.py --workload c5 --points 3000000 --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${T}_c5.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c5 3M polys: ms/step',round(d['ms_per_step'],2), d['verify'])"
done
rm -f gpurun_out/${T}_exp.jsonl
for cfg in "GPL_PIP_LD_HINTS=0" "GPL_PIP_LD_HINTS=1" "GPL_PIP_LD_HINTS=2" "GPL_PIP_LD_HINTS=3"; do
  env $cfg timeout 300 python tools/exp_pip2.py --reps 3 --tag "$cfg" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
done
cat gpurun_out/${T}_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| chk',d['checksum'],'| phases us',d['fill_phases_us'])
"
