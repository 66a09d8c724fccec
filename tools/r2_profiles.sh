#!/bin/bash
# round 2: ncu captures behind DESIGN.md section 4 — launch list of the default bench, full captures of the dominant kernel of every
# measured row (PIP stream, PIP build, hull fast / exact, LineString pairs, centroid, area, affine, WKB), at 1/10 of the BASELINE
# sizes for the op kernels (counters per launch do not need 41 GB), plus the kernel-resident timings at FULL size.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2p}
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_stream -s 3 -c 1 -o gpurun_out/${T}_pip_stream $B >> gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pip_build_fill -s 3 -c 1 -o gpurun_out/${T}_pip_build_fill $B >> gpurun_out/${T}_ncu_bench.log 2>&1
O="python tools/bench_ops.py --scale 0.1 --out gpurun_out/${T}_ops_small.json"
for k in k_hull_fast "k_hull<" k_centroid k_area k_affine k_ls_ls_fast k_wkb_fill k_wkb_write; do
  name=$(echo $k | tr -d '<')
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o gpurun_out/${T}_$name $O > gpurun_out/${T}_ncu_$name.log 2>&1
done
timeout 1500 python tools/bench_ops.py --out gpurun_out/${T}_ops_roofline.json > gpurun_out/${T}_ops.log 2>&1; tail -3 gpurun_out/${T}_ops.log
ls -la gpurun_out/${T}_*.ncu-rep
