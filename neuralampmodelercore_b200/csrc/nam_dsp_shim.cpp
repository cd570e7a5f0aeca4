// nam_dsp_shim.cpp -- the C++ face of the drop-in boundary: nam::DSP / nam::get_dsp / the activation
// switches (include/NAM/*.h) implemented on top of the C ABI (include/nam_b200.h).  Host code only; the
// arithmetic is in the CUDA kernels.  Mirrors the behaviour of reference NAM/dsp.cpp:20-201,
// NAM/get_dsp.cpp:113-181,263-273 and NAM/activations.cpp:168-232 for the calls a host makes.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/NAM/dsp.h"
#include "../../include/NAM/get_dsp.h"
#include "../../include/NAM/slimmable.h"
#include "../../include/nam_b200.h"
#include "../../include/wav.h"
#include "json_lite.h"
#include "nam_model_spec.h"

namespace
{
thread_local bool gPrewarmOnResetDefault = true;

[[noreturn]] void throw_last_error(int rc)
{
  const std::string msg = nam_b200_last_error();
  if (rc == NAM_B200_ERR_FILE)
    throw nam::NamFileValidationError(msg);
  if (rc == NAM_B200_ERR_INVALID_ARGUMENT)
    throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
} // namespace

// ---- activation switches ---------------------------------------------------------------------------
bool nam::activations::Activation::using_fast_tanh = false;

void nam::activations::Activation::enable_fast_tanh()
{
  using_fast_tanh = true;
}
void nam::activations::Activation::disable_fast_tanh()
{
  using_fast_tanh = false;
}
void nam::activations::Activation::enable_lut(std::string function_name, float, float, std::size_t)
{
  throw std::runtime_error("LUT activations (" + function_name + ") are not available on the CUDA path");
}
void nam::activations::Activation::disable_lut(std::string)
{
  // enable_lut never succeeds here, so there is nothing to restore (the reference puts the original activation back)
}

// ---- ScopedPrewarmOnResetDefault / DSP base -------------------------------------------------------
nam::ScopedPrewarmOnResetDefault::ScopedPrewarmOnResetDefault(const bool prewarmOnReset)
: mPreviousPrewarmOnReset(gPrewarmOnResetDefault)
{
  gPrewarmOnResetDefault = prewarmOnReset;
}
nam::ScopedPrewarmOnResetDefault::~ScopedPrewarmOnResetDefault()
{
  gPrewarmOnResetDefault = mPreviousPrewarmOnReset;
}

nam::DSP::DSP(const int in_channels, const int out_channels, const double expected_sample_rate)
: mExpectedSampleRate(expected_sample_rate)
, mPrewarmOnReset(gPrewarmOnResetDefault)
, mInChannels(in_channels)
, mOutChannels(out_channels)
{
  if (in_channels <= 0 || out_channels <= 0)
    throw std::runtime_error("Channel counts must be positive");
}

void nam::DSP::prewarm()
{
  if (mMaxBufferSize == 0)
    SetMaxBufferSize(NAM_DEFAULT_MAX_BUFFER_SIZE);
  const int prewarmSamples = GetPrewarmSamples();
  if (prewarmSamples == 0)
    return;
  const int bufferSize = mMaxBufferSize > 1 ? mMaxBufferSize : 1;
  std::vector<std::vector<NAM_SAMPLE>> in(mInChannels, std::vector<NAM_SAMPLE>(bufferSize, (NAM_SAMPLE)0.0));
  std::vector<std::vector<NAM_SAMPLE>> out(mOutChannels, std::vector<NAM_SAMPLE>(bufferSize, (NAM_SAMPLE)0.0));
  std::vector<NAM_SAMPLE*> inPtrs, outPtrs;
  for (auto& v : in)
    inPtrs.push_back(v.data());
  for (auto& v : out)
    outPtrs.push_back(v.data());
  for (int done = 0; done < prewarmSamples; done += bufferSize)
    this->process(inPtrs.data(), outPtrs.data(), bufferSize);
}

void nam::DSP::process(NAM_SAMPLE** input, NAM_SAMPLE** output, const int num_frames)
{
  const int shared = mInChannels < mOutChannels ? mInChannels : mOutChannels;
  for (int ch = 0; ch < shared; ch++)
    for (int i = 0; i < num_frames; i++)
      output[ch][i] = input[ch][i];
  for (int ch = shared; ch < mOutChannels; ch++)
    for (int i = 0; i < num_frames; i++)
      output[ch][i] = (NAM_SAMPLE)0.0;
}

double nam::DSP::GetLoudness() const
{
  if (!HasLoudness())
    throw std::runtime_error("Asked for loudness of a model that doesn't know how loud it is!");
  return mLoudness;
}

void nam::DSP::Reset(const double sampleRate, const int maxBufferSize)
{
  mExternalSampleRate = sampleRate;
  mHaveExternalSampleRate = true;
  SetMaxBufferSize(maxBufferSize);
  if (GetPrewarmOnReset())
    prewarm();
}

void nam::DSP::ResetAndPrewarm(const double sampleRate, const int maxBufferSize)
{
  const bool previous = GetPrewarmOnReset();
  SetPrewarmOnReset(true);
  try
  {
    Reset(sampleRate, maxBufferSize);
  }
  catch (...)
  {
    SetPrewarmOnReset(previous);
    throw;
  }
  SetPrewarmOnReset(previous);
}

void nam::DSP::SetPrewarmOnReset(const bool prewarmOnReset)
{
  mPrewarmOnReset.store(prewarmOnReset, std::memory_order_release);
}
bool nam::DSP::GetPrewarmOnReset() const
{
  return mPrewarmOnReset.load(std::memory_order_acquire);
}
void nam::DSP::SetLoudness(const double loudness)
{
  mLoudness = loudness;
  mHasLoudness = true;
}
void nam::DSP::SetMaxBufferSize(const int maxBufferSize)
{
  mMaxBufferSize = maxBufferSize;
}
double nam::DSP::GetInputLevel()
{
  return mInputLevel.level;
}
double nam::DSP::GetOutputLevel()
{
  return mOutputLevel.level;
}
bool nam::DSP::HasInputLevel()
{
  return mInputLevel.haveLevel;
}
bool nam::DSP::HasOutputLevel()
{
  return mOutputLevel.haveLevel;
}
void nam::DSP::SetInputLevel(const double inputLevel)
{
  mInputLevel.haveLevel = true;
  mInputLevel.level = (float)inputLevel;
}
void nam::DSP::SetOutputLevel(const double outputLevel)
{
  mOutputLevel.haveLevel = true;
  mOutputLevel.level = (float)outputLevel;
}

// ---- B200DSP ----------------------------------------------------------------------------------------
nam::B200DSP::B200DSP(nam_b200_model* handle, const int in_channels, const int out_channels,
                      const double expected_sample_rate, const int prewarm_samples)
: DSP(in_channels, out_channels, expected_sample_rate)
, mHandle(handle)
, mPrewarmSamples(prewarm_samples)
{
}

nam::B200DSP::~B200DSP()
{
  nam_b200_destroy(mHandle);
}

void nam::B200DSP::Reset(const double sampleRate, const int maxBufferSize)
{
  mExternalSampleRate = sampleRate;
  mHaveExternalSampleRate = true;
  SetMaxBufferSize(maxBufferSize);
  // the C ABI resets (zeroes every stream's history) and, if the handle was created with
  // prewarm_on_reset, prewarms on the device; the instance-level switch is applied here
  int rc = nam_b200_reset(mHandle, sampleRate, maxBufferSize);
  if (rc != NAM_B200_OK)
    throw_last_error(rc);
  if (GetPrewarmOnReset())
    prewarm();
}

void nam::B200DSP::prewarm()
{
  if (mMaxBufferSize == 0)
  {
    SetMaxBufferSize(NAM_DEFAULT_MAX_BUFFER_SIZE);
    const int rc0 = nam_b200_reset(mHandle, mExternalSampleRate, mMaxBufferSize);
    if (rc0 != NAM_B200_OK)
      throw_last_error(rc0);
  }
  const int rc = nam_b200_prewarm(mHandle);
  if (rc != NAM_B200_OK)
    throw_last_error(rc);
}

void nam::B200DSP::process(NAM_SAMPLE** input, NAM_SAMPLE** output, const int num_frames)
{
  // LSTM checks the fast-tanh switch at run time (reference NAM/lstm.cpp:48)
  nam_b200_set_fast_tanh(mHandle, nam::activations::Activation::using_fast_tanh ? 1 : 0);
#ifdef NAM_SAMPLE_FLOAT
  const int rc = nam_b200_process_f32_planar(mHandle, input, output, num_frames);
#else
  const int rc = nam_b200_process_f64_planar(mHandle, input, output, num_frames);
#endif
  if (rc != NAM_B200_OK)
    throw_last_error(rc); // the reference only asserts here; a CUDA failure must not pass silently
}

// ---- B200SlimmableDSP (ContainerModel::SetSlimmableSize, NAM/container.cpp:99-133) ----------------------
void nam::B200SlimmableDSP::SetSlimmableSize(const double val)
{
  const int rc = nam_b200_set_slimmable_size(mHandle, val);
  if (rc < 0)
    throw_last_error(rc);
  // the newly active sub-model was Reset; the prewarm that DSP::Reset implies is driven from here because shim
  // handles are created with prewarm_on_reset = 0 (container.cpp:117-118, dsp.cpp:130-140)
  if (rc == 1 && mPrewarmOnReset && mHaveExternalSampleRate)
  {
    const int rc2 = nam_b200_prewarm(mHandle);
    if (rc2 != NAM_B200_OK)
      throw_last_error(rc2);
  }
  nam_b200_info info;
  std::memset(&info, 0, sizeof(info));
  info.struct_size = sizeof(info);
  if (nam_b200_get_info(mHandle, &info) == NAM_B200_OK)
    mPrewarmSamples = info.prewarm_samples; // GetPrewarmSamples() follows the active sub-model
}

std::vector<double> nam::B200SlimmableDSP::GetSlimmableSizeBreakpoints() const
{
  const int n = nam_b200_slimmable_breakpoints(mHandle, nullptr, 0);
  std::vector<double> out((size_t)(n > 0 ? n : 0));
  if (n > 0)
    nam_b200_slimmable_breakpoints(mHandle, out.data(), n);
  return out;
}

// ---- get_dsp -----------------------------------------------------------------------------------------
namespace
{
std::unique_ptr<nam::DSP> wrap_handle(nam_b200_model* h)
{
  nam_b200_info info;
  std::memset(&info, 0, sizeof(info));
  info.struct_size = sizeof(info);
  const int rc = nam_b200_get_info(h, &info);
  if (rc != NAM_B200_OK)
  {
    nam_b200_destroy(h);
    throw_last_error(rc);
  }
  // a SlimmableContainer handle accepts a size (1.0 = the default, full-size sub-model: a no-op here)
  const bool slimmable = nam_b200_set_slimmable_size(h, 1.0) >= 0;
  std::unique_ptr<nam::B200DSP> dsp;
  if (slimmable)
    dsp = std::make_unique<nam::B200SlimmableDSP>(h, info.in_channels, info.out_channels, info.expected_sample_rate,
                                                  info.prewarm_samples);
  else
    dsp = std::make_unique<nam::B200DSP>(h, info.in_channels, info.out_channels, info.expected_sample_rate,
                                         info.prewarm_samples);
  // apply_metadata (reference NAM/get_dsp.cpp:205-213)
  if (info.has_loudness)
    dsp->SetLoudness(info.loudness);
  if (info.has_input_level)
    dsp->SetInputLevel(info.input_level_dbu);
  if (info.has_output_level)
    dsp->SetOutputLevel(info.output_level_dbu);
  return dsp;
}

nam_b200_options make_options(int batch)
{
  nam_b200_options o;
  nam_b200_default_options(&o);
  o.max_batch = batch;
  o.fast_tanh = nam::activations::Activation::using_fast_tanh ? 1 : 0;
  // prewarm is driven by B200DSP::Reset (instance-level switch), not by the handle
  o.prewarm_on_reset = 0;
  // compile the model when NVRTC is there (cubins are cached on disk): the plugin protocol -- one stream, 64-frame
  // process() calls -- then runs on the low-latency kernel (wavenet_lat.cuh); $NAM_B200_JIT=0 turns it off
  o.jit = 3;
  return o;
}

std::unique_ptr<nam::DSP> finish(std::unique_ptr<nam::DSP> dsp, const nam::DspLoadOptions& options, bool previous)
{
  // DspLoadOptions.prewarm overrides the default only while loading; the returned model gets the caller's
  // previous default back (reference NAM/get_dsp.cpp:263-273)
  if (options.prewarm.has_value() && dsp)
    dsp->SetPrewarmOnReset(previous);
  return dsp;
}
} // namespace

std::unique_ptr<nam::DSP> nam::get_dsp_batched(const std::filesystem::path config_filename, int batch,
                                               DspLoadOptions options)
{
  const bool previous = gPrewarmOnResetDefault;
  std::unique_ptr<ScopedPrewarmOnResetDefault> scope;
  if (options.prewarm.has_value())
    scope = std::make_unique<ScopedPrewarmOnResetDefault>(*options.prewarm);
  nam_b200_options o = make_options(batch);
  nam_b200_model* h = nullptr;
  const int rc = nam_b200_create_from_file(config_filename.string().c_str(), &o, &h);
  if (rc != NAM_B200_OK)
    throw_last_error(rc);
  return finish(wrap_handle(h), options, previous);
}

std::unique_ptr<nam::DSP> nam::get_dsp(const std::filesystem::path config_filename, DspLoadOptions options)
{
  return get_dsp_batched(config_filename, 1, options);
}

std::unique_ptr<nam::DSP> nam::get_dsp_from_json_text(const std::string& nam_json_text, DspLoadOptions options)
{
  const bool previous = gPrewarmOnResetDefault;
  std::unique_ptr<ScopedPrewarmOnResetDefault> scope;
  if (options.prewarm.has_value())
    scope = std::make_unique<ScopedPrewarmOnResetDefault>(*options.prewarm);
  nam_b200_options o = make_options(1);
  nam_b200_model* h = nullptr;
  const int rc = nam_b200_create_from_json(nam_json_text.c_str(), &o, &h);
  if (rc != NAM_B200_OK)
    throw_last_error(rc);
  return finish(wrap_handle(h), options, previous);
}

std::unique_ptr<nam::DSP> nam::get_dsp(const std::filesystem::path config_filename, dspData& returnedConfig,
                                       DspLoadOptions options)
{
  // populate_dsp_data (reference NAM/get_dsp.cpp:141-154)
  std::ifstream in(config_filename, std::ios::binary);
  if (!in.is_open())
    throw NamFileValidationError("Could not validate .nam file [" + config_filename.string()
                                 + "]: file does not exist.");
  std::stringstream ss;
  ss << in.rdbuf();
  namb200::json::Value root;
  try
  {
    root = namb200::json::Value::parse(ss.str());
  }
  catch (const namb200::json::ParseError& e)
  {
    throw NamFileValidationError("Could not parse .nam file [" + config_filename.string() + "]: " + e.what());
  }
  for (const char* key : {"version", "architecture", "config", "weights"})
    if (!root.contains(key))
      throw NamFileValidationError("Invalid .nam file [" + config_filename.string() + "]: missing required key \""
                                   + key + "\".");
  returnedConfig.version = root.at("version").as_string("version");
  verify_config_version(returnedConfig.version);
  returnedConfig.architecture = root.at("architecture").as_string("architecture");
  returnedConfig.config = root.at("config").dump();
  returnedConfig.metadata = root.get("metadata").dump();
  returnedConfig.weights.clear();
  for (const auto& w : root.at("weights").items("weights"))
    returnedConfig.weights.push_back((float)w.as_double());
  returnedConfig.expected_sample_rate = root.contains("sample_rate") ? root.at("sample_rate").as_double() : -1.0;
  return get_dsp(config_filename, options);
}

void nam::verify_config_version(const std::string version)
{
  const namb200::VersionSupport s = namb200::version_support(version);
  if (s == namb200::VersionSupport::No)
    throw std::runtime_error("Model config is an unsupported version " + version + ".");
  if (s == namb200::VersionSupport::Partial)
    std::cerr << "Model config is a partially-supported version " << version << ". Continuing with partial support."
              << std::endl;
}

// ---- wav.h -------------------------------------------------------------------------------------------
std::string dsp::wav::GetMsgForLoadReturnCode(LoadReturnCode rc)
{
  switch (rc)
  {
    case LoadReturnCode::SUCCESS: return "success";
    case LoadReturnCode::ERROR_OPENING: return "could not open the file";
    case LoadReturnCode::ERROR_NOT_RIFF: return "not a RIFF file";
    case LoadReturnCode::ERROR_NOT_WAVE: return "not a WAVE file";
    case LoadReturnCode::ERROR_MISSING_FMT: return "missing fmt chunk";
    case LoadReturnCode::ERROR_INVALID_FILE: return "invalid or truncated file";
    case LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_ALAW: return "A-law WAV files are not supported";
    case LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_MULAW: return "mu-law WAV files are not supported";
    case LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_OTHER: return "unsupported WAV format tag";
    case LoadReturnCode::ERROR_UNSUPPORTED_BITS_PER_SAMPLE: return "unsupported bits per sample";
    case LoadReturnCode::ERROR_NOT_MONO: return "only mono files are supported";
    default: return "unknown error";
  }
}

dsp::wav::LoadReturnCode dsp::wav::Load(const char* fileName, std::vector<float>& audio, double& sampleRate)
{
  std::ifstream f(fileName, std::ios::binary);
  if (!f.is_open())
    return LoadReturnCode::ERROR_OPENING;
  std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  auto u16 = [&d](size_t p) { return (unsigned)d[p] | ((unsigned)d[p + 1] << 8); };
  auto u32 = [&d](size_t p) {
    return (uint32_t)d[p] | ((uint32_t)d[p + 1] << 8) | ((uint32_t)d[p + 2] << 16) | ((uint32_t)d[p + 3] << 24);
  };
  if (d.size() < 12 || std::memcmp(d.data(), "RIFF", 4) != 0)
    return LoadReturnCode::ERROR_NOT_RIFF;
  if (std::memcmp(d.data() + 8, "WAVE", 4) != 0)
    return LoadReturnCode::ERROR_NOT_WAVE;
  size_t pos = 12, data_pos = 0, data_len = 0;
  unsigned tag = 0, channels = 0, bits = 0;
  uint32_t rate = 0;
  bool have_fmt = false;
  while (pos + 8 <= d.size())
  {
    const uint32_t size = u32(pos + 4);
    const size_t body = pos + 8;
    if (std::memcmp(d.data() + pos, "fmt ", 4) == 0 && body + 16 <= d.size())
    {
      tag = u16(body);
      channels = u16(body + 2);
      rate = u32(body + 4);
      bits = u16(body + 14);
      if (tag == 0xFFFE && size >= 26)
        tag = u16(body + 24); // WAVE_FORMAT_EXTENSIBLE: sub-format
      have_fmt = true;
    }
    else if (std::memcmp(d.data() + pos, "data", 4) == 0)
    {
      data_pos = body;
      data_len = std::min<size_t>(size, d.size() - body);
    }
    pos = body + size + (size & 1);
  }
  if (!have_fmt)
    return LoadReturnCode::ERROR_MISSING_FMT;
  if (data_pos == 0)
    return LoadReturnCode::ERROR_INVALID_FILE;
  if (tag == 6)
    return LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_ALAW;
  if (tag == 7)
    return LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_MULAW;
  if (tag != 1 && tag != 3)
    return LoadReturnCode::ERROR_UNSUPPORTED_FORMAT_OTHER;
  if (channels != 1)
    return LoadReturnCode::ERROR_NOT_MONO;
  const size_t bytes = bits / 8;
  if (bytes == 0)
    return LoadReturnCode::ERROR_UNSUPPORTED_BITS_PER_SAMPLE;
  const size_t n = data_len / bytes;
  audio.resize(n);
  const unsigned char* p = d.data() + data_pos;
  if (tag == 3 && bits == 32)
  {
    std::memcpy(audio.data(), p, n * 4);
  }
  else if (tag == 1 && bits == 16)
  {
    for (size_t i = 0; i < n; i++)
      audio[i] = (float)((int16_t)(p[2 * i] | (p[2 * i + 1] << 8)) / 32768.0);
  }
  else if (tag == 1 && bits == 24)
  {
    for (size_t i = 0; i < n; i++)
    {
      int32_t v = (int32_t)(p[3 * i] | (p[3 * i + 1] << 8) | (p[3 * i + 2] << 16));
      if (v & 0x800000)
        v -= 0x1000000;
      audio[i] = (float)(v / 8388608.0);
    }
  }
  else if (tag == 1 && bits == 32)
  {
    for (size_t i = 0; i < n; i++)
      audio[i] = (float)((int32_t)u32(data_pos + 4 * i) / 2147483648.0);
  }
  else
    return LoadReturnCode::ERROR_UNSUPPORTED_BITS_PER_SAMPLE;
  sampleRate = (double)rate;
  return LoadReturnCode::SUCCESS;
}
