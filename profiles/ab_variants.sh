#!/bin/bash
# A/B harness: build kernel variants with -D flags and time bench.py stage_ms for each (run under gpurun).
# usage: profiles/ab_variants.sh build   (here)     |    profiles/ab_variants.sh run   (on the GPU box)
cd "$(dirname "$0")/.."
CS=ss-gnn_b200/csrc
declare -A V
while IFS='=' read -r name flags; do [ -n "$name" ] && V[$name]="$flags"; done < profiles/ab_variants.txt
if [ "$1" = build ]; then
  for name in "${!V[@]}"; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
      -Iinclude -I$CS ${V[$name]} $CS/*.cu -o build_variants/lib_$name.so &
  done; wait; ls build_variants
else
  for name in $(echo "${!V[@]}" | tr ' ' '\n' | sort); do
    UGS_B200_LIB=$PWD/build_variants/lib_$name.so python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS} 2>gpurun_out/ab_err_$name.log | \
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '${V[$name]}', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()})"
  done
fi
