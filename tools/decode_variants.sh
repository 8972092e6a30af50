#!/usr/bin/env bash
# Build decode.cu variants (macro overrides) into probpose_code_b200/build/variants/lib_<name>.so; the other objects are
# the ones of the normal build.  Usage: tools/decode_variants.sh name1 "-DPP_DEC_BUDGET=60" name2 "-D..." ...
set -euo pipefail
cd "$(dirname "$0")/../probpose_code_b200"
python -m probpose_code_b200.build >/dev/null 2>&1 || (cd .. && python -m probpose_code_b200.build >/dev/null)
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
    -c csrc/decode.cu -o build/variants/decode_$name.o
  objs=$(ls build/*.o | grep -v '/decode\.o$')
  nvcc -shared -o build/variants/lib_$name.so $objs build/variants/decode_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static -ldl
  echo "built build/variants/lib_$name.so ($flags)"
done
