#!/bin/bash
# usage: tools/sweep.sh "<chunks list>" lib1.so lib2.so ...   (runs on the GPU box)
CHS="$1"; shift
for lib in "$@"; do
  for dt in fp32 bf16; do
    for ch in $CHS; do
      for clips in 8 1; do
        v=$(AFA_LIBRARY=$lib timeout 200 python bench.py --no-e2e --no-cpu-baseline --dtype $dt --clips $clips --chunks $ch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_info']['registers'], d['roofline']['kernel_info']['ctas_per_sm'], d['clocks']['sm_mhz'])")
        echo "$(basename $lib) $dt ch=$ch clips=$clips GB/s,regs,ctas,mhz: $v"
      done
    done
  done
done
