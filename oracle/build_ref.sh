#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED reference CUDA sources where they lie under
# $REF (default /root/reference) for sm_100a, plus our thin C drivers, into
# oracle/_ref/*.so (git-ignored, travels to the GPU box with gpurun).
# The reference's own build system (setup.py / CMake / torch extension) is not
# run; the only additions are
#   -include cstdint   rasterizer_impl.h:49-61 uses uint32_t/uintptr_t without
#                      <cstdint> (gcc 13 rejects it)
#   -I third_party/glm the GLM copy vendored inside each reference submodule.
# nvcc defaults (-O3 device code, fmad on, no fast-math) match what
# torch.utils.cpp_extension would have used for the reference's setup.py.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p "$OUT"
if [ ! -d "$REF/submodules" ]; then
    echo "build_ref: $REF not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi

build_raster() {  # $1 = variant tag, $2 = submodule dir, $3 = -D flag
    local tag="$1" sub="$REF/submodules/$2" def="$3"
    local so="$OUT/libref_${tag}.so"
    if [ -f "$so" ] && [ "$so" -nt "$HERE/ref_driver_raster.cu" ] && [ "${FORCE:-0}" = "0" ]; then
        echo "build_ref: $so up to date"; return
    fi
    echo "build_ref: building $so"
    nvcc -O3 $ARCH -std=c++17 -lineinfo -include cstdint -Xcompiler -fPIC -shared \
        -w -I "$sub" -I "$sub/third_party/glm" "$def" \
        "$sub/cuda_rasterizer/forward.cu" "$sub/cuda_rasterizer/backward.cu" \
        "$sub/cuda_rasterizer/rasterizer_impl.cu" "$HERE/ref_driver_raster.cu" \
        -o "$so"
}

build_raster surfel diff-surfel-rasterization -DREF_VARIANT_SURFEL
if grep -q REF_VARIANT_GAUSSIAN "$HERE/ref_driver_raster.cu"; then
    build_raster gaussian diff-gaussian-rasterization -DREF_VARIANT_GAUSSIAN
fi
if grep -q REF_VARIANT_PLANE "$HERE/ref_driver_raster.cu"; then
    build_raster plane diff-plane-rasterization -DREF_VARIANT_PLANE
fi
if [ -f "$HERE/ref_driver_filter.cu" ]; then
    so="$OUT/libref_filter.so"; sub="$REF/submodules/scaffold-filter"
    if [ ! -f "$so" ] || [ "$HERE/ref_driver_filter.cu" -nt "$so" ] || [ "${FORCE:-0}" != "0" ]; then
        echo "build_ref: building $so"
        nvcc -O3 $ARCH -std=c++17 -lineinfo -include cstdint -Xcompiler -fPIC -shared -w \
            -I "$sub" -I "$sub/third_party/glm" \
            "$sub/cuda_rasterizer/forward.cu" \
            "$sub/cuda_rasterizer/rasterizer_impl.cu" "$HERE/ref_driver_filter.cu" -o "$so"
    fi
fi
if [ -f "$HERE/ref_driver_knn.cu" ]; then
    so="$OUT/libref_knn.so"; sub="$REF/submodules/simple-knn"
    if [ ! -f "$so" ] || [ "$HERE/ref_driver_knn.cu" -nt "$so" ] || [ "${FORCE:-0}" != "0" ]; then
        echo "build_ref: building $so"
        nvcc -O3 $ARCH -std=c++17 -lineinfo -include cstdint -include cfloat -Xcompiler -fPIC -shared -w \
            -I "$sub" "$sub/simple_knn.cu" "$HERE/ref_driver_knn.cu" -o "$so"
    fi
fi
echo "build_ref: done"; ls -la "$OUT"
