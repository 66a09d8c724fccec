#!/bin/bash
# time k_pip variants on the GPU box (rebuilds the library there)
cd geopolars_b200/csrc
run() { (cd ../.. && python tools/prof_pip.py 2>&1 | grep -E "query ms|index build" | tail -2); }
build() { make -s EXTRA="$1" 2>&1 | grep -E "error"; }
for cfg in "-DGPL_PIP_MINB=3" "-DGPL_PIP_MINB=4" "-DGPL_PIP_MINB=4 -DGPL_PIP_NOALLOC=1" "-DGPL_PIP_MINB=5"; do
  touch k_pip.cu; build "$cfg"; echo "== $cfg: $(grep -A3 'k_pip_queryILi0' build/k_pip.ptxas.log | grep -E 'registers|spill' | tr '\n' ' ')"; run
done
touch k_pip.cu; build ""
