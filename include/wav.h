// wav.h -- minimal stand-in for AudioDSPTools' dsp/wav.h (an un-vendored submodule of the reference,
// used only by tools/render.cpp:12-15,131): load a PCM / IEEE-float RIFF WAVE file as mono float32.
#pragma once

#include <string>
#include <vector>

namespace dsp
{
namespace wav
{

enum class LoadReturnCode
{
  SUCCESS = 0,
  ERROR_OPENING,
  ERROR_NOT_RIFF,
  ERROR_NOT_WAVE,
  ERROR_MISSING_FMT,
  ERROR_INVALID_FILE,
  ERROR_UNSUPPORTED_FORMAT_ALAW,
  ERROR_UNSUPPORTED_FORMAT_MULAW,
  ERROR_UNSUPPORTED_FORMAT_OTHER,
  ERROR_UNSUPPORTED_BITS_PER_SAMPLE,
  ERROR_NOT_MONO,
  ERROR_OTHER
};

std::string GetMsgForLoadReturnCode(LoadReturnCode rc);

/// 16/24/32-bit PCM (scaled by 2^-(bits-1)) or 32-bit float, mono only.
LoadReturnCode Load(const char* fileName, std::vector<float>& audio, double& sampleRate);

} // namespace wav
} // namespace dsp
