#!/bin/bash
# A/B harness for ONE translation unit: profiles/ab_objs.sh build ugs_sample.cu   (here; variants from profiles/ab_variants.txt,
# "name=-DFLAG=..." per line) compiles that file once per variant and links it with the other objects of csrc/build/.
# On the GPU box: UGS_B200_LIB=$PWD/build_variants/lib_<name>.so python ...
cd "$(dirname "$0")/.."
CS=ss-gnn_b200/csrc
src=$2; base=${src%.*}
mkdir -p build_variants
python -c "import sys; sys.path.insert(0,'ss-gnn_b200'); from ugs_b200 import _lib; _lib.build()" >/dev/null
others=$(ls $CS/build/*.o | grep -v "/$base.o")
while IFS='=' read -r name flags; do
  [ -z "$name" ] && continue
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -Iinclude -I$CS \
      $flags -c $CS/$src -o build_variants/${base}_$name.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $others build_variants/${base}_$name.o -o build_variants/lib_$name.so ) &
done < profiles/ab_variants.txt
wait; ls build_variants/*.so
