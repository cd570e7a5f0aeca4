// generic_pack.h -- turn a WaveNet ModelSpec with any of the reference's options into the flat
// description (generic_desc.h) the general kernel (wavenet_generic.cuh) interprets.  Host only.
#pragma once

#include <string>
#include <vector>

#include "generic_desc.h"
#include "nam_model_spec.h"

namespace namb200
{

struct GenericPlan
{
  bool eligible = false;
  std::string why_not;
  std::vector<float> weights;
  std::vector<GLayer> layers; // main network first, then the condition_dsp sub-model's
  GNet net{}, cond{};
  bool has_cond = false;
  GConvNet convnet{}; // plan_convnet()
  long state_floats = 0; // ring storage per stream
  double macs_per_frame = 0.0; // algorithmic (grouped matrices counted at their real size)
};

/// Eligible: mono in / mono out WaveNet, at most 4 layer arrays, every per-frame vector (channels, 2 x bottleneck
/// when gated, condition, head sizes) at most 64 wide, post-stack head of at most 8 convolutions, condition_dsp
/// (if any) itself a WaveNet without a condition_dsp of its own.
GenericPlan plan_generic(const ModelSpec& ms);

/// ConvNet (NAM/convnet.cpp) for the same kernel family: at most 48 blocks, channel counts at most 64.
GenericPlan plan_convnet(const ModelSpec& ms);

} // namespace namb200
