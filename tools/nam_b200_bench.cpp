// nam_b200_bench.cpp -- C++ host tool over the C ABI: the protocol of the reference's tools/benchmodel.cpp
// (fast tanh on by default, Reset(sr, frames) incl. prewarm, timed loop of process() calls on zeros or a
// two-tone signal), extended with --batch for the batched reframing.  Usage:
//   nam_b200_bench [--no-fast-tanh] [--batch B] [--frames N] [--calls K] <model.nam>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nam_b200.h"

int main(int argc, char** argv)
{
  int batch = 1, frames = 64, calls = 1500, fast = 1;
  const char* path = nullptr;
  for (int i = 1; i < argc; i++)
  {
    std::string a(argv[i]);
    if (a == "--batch" && i + 1 < argc)
      batch = std::atoi(argv[++i]);
    else if (a == "--frames" && i + 1 < argc)
      frames = std::atoi(argv[++i]);
    else if (a == "--calls" && i + 1 < argc)
      calls = std::atoi(argv[++i]);
    else if (a == "--no-fast-tanh")
      fast = 0;
    else
      path = argv[i];
  }
  if (!path)
  {
    std::fprintf(stderr, "Usage: nam_b200_bench [--no-fast-tanh] [--batch B] [--frames N] [--calls K] <model.nam>\n");
    return 1;
  }
  nam_b200_options o;
  nam_b200_default_options(&o);
  o.max_batch = batch;
  o.fast_tanh = fast;
  nam_b200_model* m = nullptr;
  if (nam_b200_create_from_file(path, &o, &m) != NAM_B200_OK)
  {
    std::fprintf(stderr, "Failed to load model: %s\n", nam_b200_last_error());
    return 1;
  }
  nam_b200_info info;
  std::memset(&info, 0, sizeof(info));
  info.struct_size = sizeof(info);
  nam_b200_get_info(m, &info);
  if (nam_b200_reset(m, info.expected_sample_rate, frames) != NAM_B200_OK)
  {
    std::fprintf(stderr, "reset failed: %s\n", nam_b200_last_error());
    return 1;
  }
  std::vector<float> in((size_t)batch * frames), out((size_t)batch * frames);
  for (int b = 0; b < batch; b++)
    for (int i = 0; i < frames; i++)
      in[(size_t)b * frames + i] = 0.25f * std::sin(6.2831853f * 220.0f * i / 48000.0f + 0.01f * b);
  for (int w = 0; w < 3; w++)
    nam_b200_process_f32(m, in.data(), out.data(), batch, frames, frames, frames);
  const auto t1 = std::chrono::high_resolution_clock::now();
  for (int c = 0; c < calls; c++)
    if (nam_b200_process_f32(m, in.data(), out.data(), batch, frames, frames, frames) != NAM_B200_OK)
    {
      std::fprintf(stderr, "process failed: %s\n", nam_b200_last_error());
      return 1;
    }
  const auto t2 = std::chrono::high_resolution_clock::now();
  const double ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
  const double samples = (double)batch * frames * calls;
  std::printf("{\"model\": \"%s\", \"batch\": %d, \"frames\": %d, \"calls\": %d, \"fast_tanh\": %d, \"ms\": %.3f, "
              "\"msamples_per_s\": %.3f, \"rtf_48k_per_stream\": %.2f, \"last_output\": %.9g}\n",
              path, batch, frames, calls, fast, ms, samples / ms / 1e3, samples / batch / 48000.0 / (ms * 1e-3),
              (double)out[(size_t)batch * frames - 1]);
  nam_b200_destroy(m);
  return 0;
}
