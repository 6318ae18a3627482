// oracle/ref_driver_knn.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C driver around the UNMODIFIED reference simple-knn source (submodules/simple-knn/simple_knn.cu,
// compiled where it lies under /root/reference by oracle/build_ref.sh into oracle/_ref/libref_knn.so).
// Replaces only the torch glue distCUDA2 (K/spatial.cu:14-25).
// Users: tests/, tests/golden/make_golden.py, tests/gpu_bench_aux.py.  Never loaded by the product path.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_dist2_knn3(int P, float* points, float* meanDists) {
    if (P == 0) return 0;
    SimpleKNN::knn(P, (float3*)points, meanDists);
    cudaDeviceSynchronize();
    return (int)cudaGetLastError();
}
