#!/bin/bash
cd geopolars_b200/csrc
for mb in 3 4; do
  sed -i "s/#define GPL_PIP_MINB [0-9]/#define GPL_PIP_MINB $mb/" k_pip.cu
  make -s 2>&1 | grep -E "error"
  echo "== minblocks=$mb: $(grep -A3 'k_pip_queryILi0' build/k_pip.ptxas.log | grep -E 'registers|spill' | tr '\n' ' ')"
  (cd ../.. && python scratch/prof_pip.py 2>&1 | grep -E "query ms" | tail -1)
done
