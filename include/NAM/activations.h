// NAM/activations.h -- the run-time switches of the reference's activation registry that callers touch
// (reference NAM/activations.h:163-167, NAM/activations.cpp:168-232).  The arithmetic itself lives in the
// CUDA kernels (neuralampmodelercore_b200/csrc/wavenet_fused.cuh); here only the process-wide flags exist.
#pragma once

#include <cstddef>
#include <string>

namespace nam
{
namespace activations
{

class Activation
{
public:
  /// Replace "Tanh" by the rational fast_tanh for models loaded AFTERWARDS (WaveNet captures the choice at
  /// construction, LSTM reads the flag at run time) -- same semantics as the reference.
  static void enable_fast_tanh();
  static void disable_fast_tanh();
  static bool using_fast_tanh;
  /// The reference can swap Tanh/Sigmoid/SiLU for an interpolated lookup table.  The CUDA path has no LUT
  /// mode (it would be slower than the arithmetic on a GPU): these throw std::runtime_error.
  static void enable_lut(std::string function_name, float min, float max, std::size_t n_points);
  static void disable_lut(std::string function_name);
};

} // namespace activations
} // namespace nam
