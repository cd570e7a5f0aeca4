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
  /// The reference can swap Tanh/Sigmoid/SiLU for an interpolated lookup table (NAM/activations.h:371-422), which CHANGES
  /// the arithmetic.  The CUDA path has no LUT mode (a table lookup is slower than the arithmetic on a GPU, and silently
  /// computing the exact function instead would not be the reference's output): enable_lut throws std::runtime_error,
  /// disable_lut is a no-op (tests/test_host_logic.py::test_lut_switch_is_refused_loudly).
  static void enable_lut(std::string function_name, float min, float max, std::size_t n_points);
  static void disable_lut(std::string function_name);
};

} // namespace activations
} // namespace nam
