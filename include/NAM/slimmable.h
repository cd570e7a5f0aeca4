// NAM/slimmable.h -- interface of models that can trade quality for cost at run time (reference
// NAM/slimmable.h:13-29).  Kept so hosts that dynamic_cast<nam::SlimmableModel*> (tools/benchmodel.cpp:93,
// tools/render.cpp:119) compile; the B200 models do not implement it (slimming is control-plane, out of
// scope of the hot path), so the cast yields nullptr and the tools report that.
#pragma once

#include <vector>

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

} // namespace nam
