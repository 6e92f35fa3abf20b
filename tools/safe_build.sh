#!/bin/bash
# Rebuilds the native libraries into temporary names and renames them into place (atomic on one filesystem), so that a
# gpurun snapshot taken at any moment sees a complete library.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" \
     -o robust_cvd_b200/.librcvd_b200.so.tmp robust_cvd_b200/csrc/rcvd_api.cu 2>&1 | grep -v "Remark\|warning #\|^$\|\^\|detected during\|for (int i\|instantiation" || true
mv robust_cvd_b200/.librcvd_b200.so.tmp robust_cvd_b200/librcvd_b200.so
EXT=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
make -s -C robust_cvd_b200/host OUT=.lib_python.tmp
mv robust_cvd_b200/host/.lib_python.tmp robust_cvd_b200/host/lib_python$EXT
make -s -C oracle
echo built
