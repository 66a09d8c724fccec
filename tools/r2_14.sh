#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2p}
timeout 900 python -m pytest tests/test_gpu_pip.py -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
for w in c2 c4; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$w.json')); print('$w value %.3g ms/step %.3f kernel %.4f call %.4f frac %.4f'%(d['value'],d['ms_per_step'],d['config']['kernel_ms'],d['config']['op_call_ms'],d['roofline']['frac']))" || tail -5 gpurun_out/${T}_bench_$w.err
done
