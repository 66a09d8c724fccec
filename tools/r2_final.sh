#!/bin/bash
# round-2 final pass on one B200: the whole GPU suite, the bench lines of every arm, the reference arm, the launch list and the
# ncu captures of the kernels that changed last (PIP stream / build, hull), the op table at full size.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2f}
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_c2_reference.json 2> gpurun_out/${T}_bench_c2_reference.err; echo "bench ref rc=$?"
for w in c4 c3 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "bench $w rc=$?"
done
python - <<PY
import json
for w in ("c2","c2_reference","c4","c3","c5"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%w))
        print(w, "value %.4g ms/step %.3f"%(d["value"], d["ms_per_step"]), "kernel_ms", d.get("config",{}).get("kernel_ms"), "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(w, "no json", e)
PY
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_stream -s 3 -c 1 -f -o gpurun_out/${T}_pip_stream $B >> gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_build_fill -s 3 -c 1 -f -o gpurun_out/${T}_pip_build_fill $B >> gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_deferred -s 3 -c 1 -f -o gpurun_out/${T}_pip_deferred $B >> gpurun_out/${T}_ncu_bench.log 2>&1
O="python tools/bench_ops.py --scale 0.1 --out gpurun_out/${T}_ops_small.json"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_hull_fast" -s 1 -c 1 -f -o gpurun_out/${T}_hull_fast $O > gpurun_out/${T}_ncu_hull_fast.log 2>&1
timeout 1500 python tools/bench_ops.py --out gpurun_out/${T}_ops_roofline.json > gpurun_out/${T}_ops.log 2>&1; tail -3 gpurun_out/${T}_ops.log
ls -la gpurun_out/${T}_*.ncu-rep
