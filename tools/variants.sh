#!/bin/bash
# usage: variants.sh "<name>:<nvcc -D flags>" ...   build each variant of libbslam.so here (CPU box), then
# run `gpurun -- bash tools/variants.sh run` to time them all on the GPU.
if [ "$1" == "run" ]; then
  for so in build/variants/*.so; do echo "== $so"; BSLAM_LIB=$PWD/$so timeout 300 python tools/time_phases.py 2>&1 | tail -1; done
  exit 0
fi
mkdir -p build/variants; rm -f build/variants/*.so
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $flags -shared -o build/variants/$name.so pyslam_b200/csrc/solver.cu &
done
wait; ls -la build/variants
