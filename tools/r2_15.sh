#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_pip.py tests/test_gpu_join.py tests/test_gpu_fullsize.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
rm -f gpurun_out/${T}_exp.jsonl
timeout 300 python tools/exp_pip2.py --reps 3 --tag "c2" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
timeout 300 python tools/exp_pip2.py --reps 3 --polys 1000 --grid 32 --cell 31.25 --points 125000000 --tag "c4" >> gpurun_out/${T}_exp.jsonl 2>> gpurun_out/${T}_exp.err
cat gpurun_out/${T}_exp.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'],'| build',round(d['build_ms_min'],3),'| query',round(d['query_ms_min'],3),'| chk',d['checksum'],'| phases us',d['fill_phases_us'])
"
for w in c2 c4; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$w.json')); print('$w value %.4g ms/step %.3f kernel %.4f call %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['config']['kernel_ms'],d['config']['op_call_ms'],d['roofline']['frac']))" || tail -5 gpurun_out/${T}_bench_$w.err
done
