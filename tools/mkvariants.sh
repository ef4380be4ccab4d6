#!/bin/bash
# tools/mkvariants.sh name1:"-DFOO=1 -DBAR=2" name2:"..." — build build/rt_<name>.so variants of librt_core.so (A/B tuning; see tools/ab.sh)
mkdir -p build
cd dxrexperiments_b200/csrc
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC $flags -shared -o ../../build/rt_$name.so api.cu build.cu accel_ops.cu pipeline.cu denoise.cu comm.cu -ldl 2> ../../build/rt_$name.log || echo "FAILED $name" ) &
done
wait
ls -la ../../build/*.so
