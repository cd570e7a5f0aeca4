// Host-side launchers of the tensor-core WaveNet kernel (wavenet_tc.cuh), in a translation unit of their own so that the
// kernel is an OPTION of the build: `NAM_B200_BUILD_TC=1 python -m neuralampmodelercore_b200._build` compiles its six
// instantiations; the default build carries stubs (the kernel is slower than the FP32 path on every model of the reference's
// families -- DESIGN.md section 2.2 -- and cost a minute of compile time and 1.3 MB of library).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include "wavenet_desc.h"
#include "wavenet_pack.h"

namespace namb200
{
bool tc_built(); // was the library built with the tensor-core kernel?
size_t tc_smem_bytes(const WaveNetPlan& plan); // dynamic shared memory of one CTA (0 when not built)
int tc_occupancy(int c0, int c1, size_t smem); // resident CTAs per SM
// throws std::runtime_error (not built / no instantiation for the channel pair / launch error)
void tc_launch(int c0, int c1, const WaveNetKernelParams& kp, int image_float4, int n_layers, int grid, size_t smem, cudaStream_t st);
} // namespace namb200
