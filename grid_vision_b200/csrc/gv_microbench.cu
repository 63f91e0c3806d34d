// gv_microbench.cu — measured atomic-throughput ceilings for K2 (binning) and K3 (raycast).
//
// SURVEY §8.d asks for the raycast's cells/s to be read against "atomicAdd int32 to shared,
// conflict-free, and RED.global to L2-resident lines" measured on the same device.  These are not
// part of the hot path: bench.py calls gv_microbench_atomics once and reports k_sweep_walk's and
// k_points' atomic rates as fractions of the matching pattern.
#include "gridvision_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ unsigned mix(unsigned v)
{
  v ^= v >> 16; v *= 0x85EBCA6Bu;
  v ^= v >> 13; v *= 0xC2B2AE35u;
  v ^= v >> 16;
  return v;
}

// pattern: where the 32 lanes of a warp aim in one operation
//   0 spread     : 32 independent random cells                         (binning, far field)
//   1 contiguous : 32 consecutive cells at a random base               (raycast, y-major batches)
//   2 strided    : 32 cells `stride` apart at a random base            (raycast, x-major batches)
//   3 runs4      : 8 random cells, 4 consecutive lanes share each      (binning, near field)
//   4 same       : one random cell for the whole warp                  (worst-case contention)
__device__ __forceinline__ size_t aim(int pattern, unsigned warp_op, unsigned lane, unsigned mask, unsigned stride)
{
  switch (pattern) {
  case 0: return mix(warp_op * 32u + lane) & mask;
  case 1: return ((mix(warp_op) & mask) & ~31u) + lane;
  case 2: return ((size_t)(mix(warp_op) & mask) + (size_t)lane * stride) & mask;
  case 3: return mix(warp_op * 8u + (lane >> 2)) & mask;
  default: return mix(warp_op) & mask;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_mb_red(T *plane, unsigned mask, unsigned stride, int pattern, int reps)
{
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#pragma unroll 1
  for (int r = 0; r < reps; ++r) atomicAdd(plane + aim(pattern, warp * (unsigned)reps + r, lane, mask, stride), (T)1);
}

// conflict-free shared-memory atomics: lane l of every warp owns bank l
__global__ void __launch_bounds__(256) k_mb_atoms(int *sink, int reps)
{
  __shared__ int s[256 * 4];
  for (int i = threadIdx.x; i < 1024; i += 256) s[i] = 0;
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) atomicAdd(&s[((r & 3) << 8) + threadIdx.x], 1);
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = s[0] + s[1023];
}

}  // namespace

extern "C" GV_API int gv_microbench_atomics(int device, int kind, int pattern, size_t ncells_pow2, unsigned stride,
                                     int reps, double *ops_per_s_out)
{
  if (!ops_per_s_out || reps < 1 || ncells_pow2 < 1024 || (ncells_pow2 & (ncells_pow2 - 1)) != 0 ||
      ncells_pow2 > (1ull << 31))
    return GV_ERR_INVALID;
  *ops_per_s_out = 0.0;
  if (cudaSetDevice(device) != cudaSuccess) return GV_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return GV_ERR_NO_DEVICE;
  const unsigned blocks = (unsigned)prop.multiProcessorCount * 8u;
  void *plane = nullptr;
  const size_t bytes = ncells_pow2 * 8;
  if (cudaMalloc(&plane, bytes) != cudaSuccess) return GV_ERR_CUDA;
  cudaMemset(plane, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned mask = (unsigned)(ncells_pow2 - 1);
  auto launch = [&]() {
    if (kind == 0) k_mb_red<unsigned long long><<<blocks, 256>>>((unsigned long long *)plane, mask, stride, pattern, reps);
    else if (kind == 1) k_mb_red<int><<<blocks, 256>>>((int *)plane, mask, stride, pattern, reps);
    else k_mb_atoms<<<blocks, 256>>>((int *)plane, reps);
  };
  launch();  // warm-up: the plane becomes L2-resident when it fits
  cudaEventRecord(e0);
  const int iters = 5;
  for (int i = 0; i < iters; ++i) launch();
  cudaEventRecord(e1);
  int rc = GV_OK;
  if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = GV_ERR_CUDA;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (rc == GV_OK && ms > 0.f) *ops_per_s_out = (double)blocks * 256.0 * reps * iters / (ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(plane);
  return rc;
}
