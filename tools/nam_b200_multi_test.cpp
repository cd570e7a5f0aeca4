// nam_b200_multi_test.cpp -- a C++ host driving several GPUs through the C ABI alone (include/nam_b200.h), no Python, no
// CUDA headers: one nam_b200_multi handle over the given devices against one single-device handle on the first device.
// The shards must reproduce the single handle bit for bit (streams are independent: BASELINE.json config 5 is a shard).
//   usage: nam_b200_multi_test model.nam <streams> <frames> <device> [<device> ...]     (a device may repeat)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nam_b200.h"

static int die(const char* what)
{
  std::fprintf(stderr, "%s: %s\n", what, nam_b200_last_error());
  return 1;
}

int main(int argc, char** argv)
{
  if (argc < 5)
  {
    std::fprintf(stderr, "usage: %s model.nam streams frames device [device ...]\n", argv[0]);
    return 2;
  }
  const char* path = argv[1];
  const int B = std::atoi(argv[2]), n = std::atoi(argv[3]);
  std::vector<int32_t> devices;
  for (int i = 4; i < argc; i++)
    devices.push_back(std::atoi(argv[i]));

  nam_b200_options o;
  nam_b200_default_options(&o);
  o.max_batch = B;
  o.fast_tanh = 1;
  nam_b200_multi* mm = nullptr;
  if (nam_b200_multi_create_from_file(path, &o, devices.data(), (int)devices.size(), &mm))
    return die("multi_create");
  o.device = devices[0];
  nam_b200_model* single = nullptr;
  if (nam_b200_create_from_file(path, &o, &single))
    return die("create");
  if (nam_b200_multi_reset(mm, 48000.0, n) || nam_b200_reset(single, 48000.0, n))
    return die("reset");

  std::vector<float> x((size_t)B * n), y_multi(x.size()), y_single(x.size());
  for (int b = 0; b < B; b++)
    for (int t = 0; t < n; t++)
      x[(size_t)b * n + t] = (0.5f + 0.5f * b / B) * (0.25f * std::sin(6.2831853f * 220.0f * t / 48000.0f + 0.1f * b)
                                                      + 0.10f * std::sin(6.2831853f * 1230.0f * t / 48000.0f));
  // pinned once through the library: the copies then overlap the kernels
  if (nam_b200_pin_host_buffer(x.data(), (int64_t)x.size() * 4) || nam_b200_pin_host_buffer(y_multi.data(), (int64_t)x.size() * 4))
    return die("pin");
  if (nam_b200_host_buffer_is_pinned(x.data()) != 1 || nam_b200_host_buffer_is_pinned(y_single.data()) != 0)
    return die("is_pinned");

  double best = 1e30;
  for (int rep = 0; rep < 3; rep++) // state carries over: three consecutive calls, compared call by call
  {
    const auto t0 = std::chrono::steady_clock::now();
    if (nam_b200_multi_process_f32(mm, x.data(), y_multi.data(), B, n, n, n))
      return die("multi_process");
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt < best)
      best = dt;
    if (nam_b200_process_f32(single, x.data(), y_single.data(), B, n, n, n)) // pageable output: allowed, just slower
      return die("process");
    if (std::memcmp(y_multi.data(), y_single.data(), x.size() * 4) != 0)
    {
      std::fprintf(stderr, "call %d: the sharded result differs from the single-device result\n", rep);
      return 3;
    }
  }
  for (int i = 0; i < nam_b200_multi_device_count(mm); i++)
  {
    int32_t dev = -1, first = -1, cnt = -1;
    nam_b200_multi_shard(mm, i, &dev, &first, &cnt);
    std::printf("part %d: device %d, streams [%d, %d)\n", i, dev, first, first + cnt);
  }
  std::printf("OK %d devices, %d streams x %d frames, bit-identical to one device; %.1f Msamples/s end to end (host buffers)\n",
              (int)devices.size(), B, n, (double)B * n / best / 1e6);
  nam_b200_unpin_host_buffer(x.data());
  nam_b200_unpin_host_buffer(y_multi.data());
  nam_b200_multi_destroy(mm);
  nam_b200_destroy(single);
  return 0;
}
