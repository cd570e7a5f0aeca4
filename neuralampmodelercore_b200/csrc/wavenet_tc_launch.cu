// see wavenet_tc_launch.h
#include "wavenet_tc_launch.h"

#include <algorithm>
#include <stdexcept>
#include <string>

#ifndef NAM_B200_WITH_TC
#define NAM_B200_WITH_TC 0
#endif

#if NAM_B200_WITH_TC
#include "wavenet_tc.cuh"

namespace namb200
{
namespace
{
void check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess)
    throw std::runtime_error(std::string(what) + " failed: " + cudaGetErrorString(e));
}

template <int C0, int C1>
void launch_variant(const WaveNetKernelParams& kp, int image_float4, int n_layers, int grid, size_t smem, cudaStream_t st)
{
  auto kern = wavenet_tc_kernel<C0, C1>;
  // the kernel also has a few bytes of static shared memory (mbarriers), so ask for what is needed, not the maximum
  check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(tc kernel)");
  check(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
        "cudaFuncSetAttribute(tc kernel carve-out)");
  kern<<<grid, kTcThreads, smem, st>>>(kp, image_float4, n_layers);
  check(cudaGetLastError(), "tensor-core kernel launch");
}

template <int C0, int C1>
int occupancy_variant(size_t smem)
{
  auto kern = wavenet_tc_kernel<C0, C1>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  // shared memory decides (227 KB per SM, 1 KB reserved per CTA); registers allow 3 (launch bounds)
  const int by_smem = (int)((227 * 1024) / (smem + 1024 + 64));
  return std::max(1, std::min(3, by_smem));
}

#define TC_DISPATCH(FN, ...)                                                                                         \
  switch (c0 * 100 + c1)                                                                                             \
  {                                                                                                                  \
    case 800: return FN<8, 0>(__VA_ARGS__);                                                                          \
    case 1600: return FN<16, 0>(__VA_ARGS__);                                                                        \
    case 808: return FN<8, 8>(__VA_ARGS__);                                                                          \
    case 816: return FN<8, 16>(__VA_ARGS__);                                                                         \
    case 1608: return FN<16, 8>(__VA_ARGS__);                                                                        \
    case 1616: return FN<16, 16>(__VA_ARGS__);                                                                       \
    default: throw std::runtime_error("no tensor-core WaveNet kernel for channel pair " + std::to_string(c0) + "/"   \
                                      + std::to_string(c1));                                                         \
  }
} // namespace

bool tc_built() { return true; }

size_t tc_smem_bytes(const WaveNetPlan& plan)
{
  const int cmax = std::max(plan.cp[0], plan.cp[1]);
  const size_t pm = cmax / 4;
  const size_t wimg4 = (size_t)(plan.tc_max_image_floats + 3) / 4;
  return (2 * wimg4 + 2 * pm * kTcTW + 4 * pm * kTcM) * 16;
}

int tc_occupancy(int c0, int c1, size_t smem) { TC_DISPATCH(occupancy_variant, smem) }

void tc_launch(int c0, int c1, const WaveNetKernelParams& kp, int image_float4, int n_layers, int grid, size_t smem, cudaStream_t st)
{
  TC_DISPATCH(launch_variant, kp, image_float4, n_layers, grid, smem, st)
}
} // namespace namb200

#else // ---- built without the tensor-core kernel -----------------------------------------------------------------------

namespace namb200
{
bool tc_built() { return false; }
size_t tc_smem_bytes(const WaveNetPlan&) { return 0; }
int tc_occupancy(int, int, size_t) { return 0; }
void tc_launch(int, int, const WaveNetKernelParams&, int, int, int, size_t, cudaStream_t)
{
  throw std::runtime_error("this library was built without the tensor-core kernel");
}
} // namespace namb200
#endif
