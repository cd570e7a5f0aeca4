// NAM/slimmable.h -- interface of models that can trade quality for cost at run time (reference
// NAM/slimmable.h:13-29).  Kept so hosts that dynamic_cast<nam::SlimmableModel*> (tools/benchmodel.cpp:93,
// tools/render.cpp:119) work: a "SlimmableContainer" file (NAM/container.cpp) loads as a B200SlimmableDSP,
// every other model as a plain B200DSP (the cast yields nullptr and the tools report that, as in the reference).
#pragma once

#include <vector>

#include "dsp.h"

namespace nam
{

class SlimmableModel
{
public:
  virtual ~SlimmableModel() = default;
  /// 0.0 = smallest, 1.0 = full size
  virtual void SetSlimmableSize(const double val) = 0;
  virtual std::vector<double> GetSlimmableSizeBreakpoints() const { return {}; }
};

/// ContainerModel of the reference (NAM/container.h:29): the handle carries one complete sub-model per size and
/// serves every call from the active one; SetSlimmableSize switches (nam_b200_set_slimmable_size).
class B200SlimmableDSP : public B200DSP, public SlimmableModel
{
public:
  using B200DSP::B200DSP;
  void SetSlimmableSize(const double val) override;
  std::vector<double> GetSlimmableSizeBreakpoints() const override;
};

} // namespace nam
