#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2l}
export GPL_HULL_FUSED=0
echo skip tests
tail -4 gpurun_out/${T}_pytest.log
rm -f gpurun_out/${T}_exp.jsonl
for cfg in "GPL_PIP_LD_HINTS=1" "GPL_PIP_LD_HINTS=5" "GPL_PIP_LD_HINTS=3"; do
  env $cfg timeout 300 python tools/exp_pip2.py --reps 3 --tag "c2 $cfg" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
done
for cfg in "GPL_PIP_LD_HINTS=1" "GPL_PIP_LD_HINTS=5"; do
  env $cfg timeout 300 python tools/exp_pip2.py --reps 3 --polys 1000 --grid 32 --cell 31.25 --points 125000000 --tag "c4 $cfg" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
done
cat gpurun_out/${T}_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| chk',d['checksum'],'| phases us',d['fill_phases_us'])
"
