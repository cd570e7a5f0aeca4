// ref_wrapper.cpp -- a C entry to the UNMODIFIED reference (NeuralAmpModelerCore), compiled from its sources
// where they lie under /root/reference against oracle/eigen_shim (see eigen_shim/Eigen/Dense for why).
// TEST INFRASTRUCTURE: only tests/ and bench.py's CPU arms may load the resulting oracle/_ref/libnam_ref.so.
// Everything below goes through the reference's public API: nam::get_dsp, DSP::Reset / process / prewarm,
// Activation::enable_fast_tanh, nam::SlimmableModel (NAM/get_dsp.h, NAM/dsp.h, NAM/activations.h, NAM/slimmable.h).
#include <cstring>
#include <filesystem>
#include <memory>
#include <string>
#include <vector>

#include "NAM/activations.h"
#include "NAM/dsp.h"
#include "NAM/get_dsp.h"
#include "NAM/slimmable.h"

namespace
{
thread_local std::string g_err;
struct Handle
{
  std::unique_ptr<nam::DSP> dsp;
  std::vector<NAM_SAMPLE> in, out;
};
} // namespace

// both variants of the library are loaded into one test process: keep every reference symbol local to its .so
#pragma GCC visibility push(default)
extern "C" {

const char* namref_last_error(void)
{
  return g_err.c_str();
}

// The fast-tanh switch is process-global in the reference (activations.cpp:168-177) and is read when a model is
// built: set it, load, and leave it set for the model's lifetime (the LSTM reads it at run time too).
void* namref_create(const char* nam_path, int fast_tanh)
{
  try
  {
    if (fast_tanh)
      nam::activations::Activation::enable_fast_tanh();
    else
      nam::activations::Activation::disable_fast_tanh();
    auto h = std::make_unique<Handle>();
    h->dsp = nam::get_dsp(std::filesystem::path(nam_path));
    if (!h->dsp)
    {
      g_err = "get_dsp returned null";
      return nullptr;
    }
    return h.release();
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return nullptr;
  }
}

void namref_destroy(void* p)
{
  delete static_cast<Handle*>(p);
}

int namref_reset(void* p, double sample_rate, int max_buffer_size)
{
  try
  {
    static_cast<Handle*>(p)->dsp->Reset(sample_rate, max_buffer_size);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// DSP::prewarm() on its own: continues from the instance's current state (NAM/dsp.cpp:67-101)
int namref_prewarm(void* p)
{
  try
  {
    static_cast<Handle*>(p)->dsp->prewarm();
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

int namref_prewarm_samples(void* p)
{
  return static_cast<Handle*>(p)->dsp->GetPrewarmSamples();
}
int namref_in_channels(void* p)
{
  return static_cast<Handle*>(p)->dsp->NumInputChannels();
}
int namref_out_channels(void* p)
{
  return static_cast<Handle*>(p)->dsp->NumOutputChannels();
}
double namref_expected_sample_rate(void* p)
{
  return static_cast<Handle*>(p)->dsp->GetExpectedSampleRate();
}

// mono float in / float out through DSP::process(NAM_SAMPLE**, NAM_SAMPLE**, n); n <= max_buffer_size
int namref_process_f32(void* p, const float* in, float* out, int n)
{
  Handle* h = static_cast<Handle*>(p);
  try
  {
    h->in.assign(in, in + n);
    h->out.assign((size_t)n, 0);
    NAM_SAMPLE* ip = h->in.data();
    NAM_SAMPLE* op = h->out.data();
    h->dsp->process(&ip, &op, n);
    for (int i = 0; i < n; i++)
      out[i] = (float)h->out[(size_t)i];
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// multi-channel: in = [in_channels][n] planes, out = [out_channels][n] planes, through the same DSP::process
int namref_process_planar_f32(void* p, const float* in, float* out, int n)
{
  Handle* h = static_cast<Handle*>(p);
  try
  {
    const int ci = h->dsp->NumInputChannels(), co = h->dsp->NumOutputChannels();
    std::vector<std::vector<NAM_SAMPLE>> xi((size_t)ci), xo((size_t)co);
    std::vector<NAM_SAMPLE*> ip((size_t)ci), op((size_t)co);
    for (int c = 0; c < ci; c++)
    {
      xi[(size_t)c].assign(in + (size_t)c * n, in + (size_t)(c + 1) * n);
      ip[(size_t)c] = xi[(size_t)c].data();
    }
    for (int c = 0; c < co; c++)
    {
      xo[(size_t)c].assign((size_t)n, 0);
      op[(size_t)c] = xo[(size_t)c].data();
    }
    h->dsp->process(ip.data(), op.data(), n);
    for (int c = 0; c < co; c++)
      for (int i = 0; i < n; i++)
        out[(size_t)c * n + i] = (float)xo[(size_t)c][(size_t)i];
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

// whole signal in blocks (the tools' protocol: tools/render.cpp:147-176, tools/benchmodel.cpp:128-133)
int namref_run_f32(void* p, const float* in, float* out, long n_total, int block)
{
  for (long pos = 0; pos < n_total; pos += block)
  {
    const int n = (int)((n_total - pos) < block ? (n_total - pos) : block);
    const int rc = namref_process_f32(p, in + pos, out + pos, n);
    if (rc != 0)
      return rc;
  }
  return 0;
}

int namref_set_slimmable_size(void* p, double value)
{
  auto* s = dynamic_cast<nam::SlimmableModel*>(static_cast<Handle*>(p)->dsp.get());
  if (!s)
  {
    g_err = "model does not implement nam::SlimmableModel";
    return -1;
  }
  try
  {
    s->SetSlimmableSize(value);
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return -1;
  }
}

} // extern "C"
#pragma GCC visibility pop
