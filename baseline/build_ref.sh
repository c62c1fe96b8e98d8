#!/bin/bash
# Builds the UNMODIFIED reference (projectchrono/DEM-Engine) and baseline/run_ref.cpp against it.
#   baseline/_ref/build      the reference's own cmake build tree (git-ignored; travels to the GPU box with gpurun).  It
#                            must live at the path it runs from: the kernel / data directory is baked in at configure
#                            time (src/core/utils/RuntimeData.cpp.in of the reference) and /root/repo on the GPU box is
#                            a link to the shipped copy, so /root/repo/baseline/_ref/build is valid in both places.
#   baseline/_ref/run_ref    the driver program
# Nothing of the reference's sources is copied into the repository; the build reads them from /root/reference.
set -e
REF=${REF:-/root/reference}
OUT=/root/repo/baseline/_ref
if [ ! -d "$REF" ]; then echo "build_ref.sh: $REF not present (GPU box): using the prebuilt files"; exit 0; fi
mkdir -p "$OUT"
if [ ! -f "$OUT/build/libsimulator_multi_gpu.a" ]; then
  cmake -G Ninja -S "$REF" -B "$OUT/build" -DCMAKE_BUILD_TYPE=Release \
        -DCUB_DIR=/usr/local/cuda/lib64/cmake/cub -Dlibcudacxx_DIR=/usr/local/cuda/lib64/cmake/libcudacxx \
        -DCMAKE_CUDA_ARCHITECTURES=100 > "$OUT/cmake_config.log" 2>&1
  ninja -C "$OUT/build" -j"${JOBS:-6}" simulator_multi_gpu > "$OUT/ninja_build.log" 2>&1
fi
CUDA=/usr/local/cuda
g++ -O2 -std=gnu++17 -DDEME_BEING_CMAKE_COMPILED -I"$REF/src" -I"$OUT/build/src" -isystem $CUDA/targets/x86_64-linux/include \
    -o "$OUT/run_ref" /root/repo/baseline/run_ref.cpp \
    -Wl,-rpath,$CUDA/targets/x86_64-linux/lib:"$OUT/build/src/core" \
    "$OUT/build/libsimulator_multi_gpu.a" $CUDA/targets/x86_64-linux/lib/libcudart.so $CUDA/targets/x86_64-linux/lib/libnvrtc.so \
    -L$CUDA/targets/x86_64-linux/lib/stubs -lcuda -ldl "$OUT/build/src/core/libDEMERuntimeDataHelper.so" \
    -L$CUDA/targets/x86_64-linux/lib -lcudadevrt -lcudart_static -lrt -lpthread -ldl
echo "built $OUT/run_ref"
