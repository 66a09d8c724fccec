#!/bin/bash
# multi-GPU bench arms: bash tools/r2_multi.sh N [workloads...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}; shift
W=${@:-"c2 c4 c5 c3"}
nvidia-smi topo -m > gpurun_out/r2m_N${N}_topo.txt 2>&1
for w in $W; do
  steps=5; [ "$w" = "c2" ] && steps=10; [ "$w" = "c5" ] && steps=3; [ "$w" = "c3" ] && steps=3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --workload $w --steps $steps --warmup 3 --no-cpu > gpurun_out/r2m_N${N}_$w.json 2> gpurun_out/r2m_N${N}_$w.err
  echo "== N=$N $w rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2m_N${N}_$w.json')); print('value %.4g ms/step %.3f kernel_ms %.3f e2e %.4g verify %s'%(d['value'],d['ms_per_step'],d['config']['kernel_ms'],d['e2e']['value'],d.get('verify')))
except Exception as e:
    print('no json:',e); print(open('gpurun_out/r2m_N${N}_$w.err').read()[-1500:])
"
done
