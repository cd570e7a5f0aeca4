// tc_probe.cu -- stage-1 experiment for a tcgen05 formulation of the WaveNet layer.
// Validates, on one CTA, the pieces the real kernel would need:
//   1. tcgen05.mma kind::tf32, M=128 (time steps) x N=16 (out channels) x K=8 (in channels), A and B from
//      shared memory in the K-major NO-SWIZZLE canonical layout, D in tensor memory, read back with
//      tcgen05.ld 32x32b (TMEM lane == thread == time step);
//   2. a dilated tap as a SHIFTED descriptor: A's start address moved by an arbitrary number of 16-byte
//      rows (not a multiple of the 8-row core matrix);
//   3. accumulation of several MMAs (taps x channel halves) into one accumulator;
//   4. accuracy of the 3xTF32 split (hi*hi + lo*hi + hi*lo) against an fp64 reference, vs plain fp32 FMA.
// Prints one JSON object per test.  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe tools/tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                        \
  do                                                                                                                 \
  {                                                                                                                  \
    cudaError_t e = (x);                                                                                             \
    if (e != cudaSuccess)                                                                                            \
    {                                                                                                                \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                                                        \
      exit(1);                                                                                                       \
    }                                                                                                                \
  } while (0)

constexpr int TW = 320; // columns (time steps) per 4-channel plane, like the real kernel's tile (halo 64 + 256)
constexpr int HALO = 64;
constexpr int C = 16; // channels
constexpr int NTAP = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, no swizzle, version 1 (sm_100): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | 1<<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32)
         | (1ull << 46);
}

// c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
    "}\n" ::"r"(tmem_d),
    "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ float tf32_hi(float x)
{
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); // truncate to 10 mantissa bits
}

// mode 0: plain TF32 (operands truncated by hardware); mode 1: 3xTF32 split
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ h /* [C][TW] rows = channel */,
                                                    const float* __restrict__ w /* [NTAP][C out][C in] */,
                                                    float* __restrict__ out /* [128][C] */, int dilation, int mode,
                                                    int t0 /* first time step (column index in the tile, >= 2*dilation) */)
{
  // planes of 4 channels: A_hi / A_lo [C/4][TW] float4; B_hi / B_lo [NTAP][C/4 in-chunks][16 out][4]
  __shared__ __align__(128) float4 a_hi[(C / 4) * TW];
  __shared__ __align__(128) float4 a_lo[(C / 4) * TW];
  __shared__ __align__(128) float4 b_hi[NTAP * (C / 4) * C];
  __shared__ __align__(128) float4 b_lo[NTAP * (C / 4) * C];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x;

  for (int idx = tid; idx < (C / 4) * TW; idx += 128)
  {
    const int pl = idx / TW, col = idx % TW;
    float v[4], hi[4], lo[4];
    for (int i = 0; i < 4; i++)
    {
      v[i] = h[(pl * 4 + i) * TW + col];
      hi[i] = tf32_hi(v[i]);
      lo[i] = v[i] - hi[i];
    }
    a_hi[idx] = make_float4(mode ? hi[0] : v[0], mode ? hi[1] : v[1], mode ? hi[2] : v[2], mode ? hi[3] : v[3]);
    a_lo[idx] = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
  for (int idx = tid; idx < NTAP * (C / 4) * C; idx += 128)
  {
    const int k = idx / ((C / 4) * C), rem = idx % ((C / 4) * C), chunk = rem / C, o = rem % C;
    float v[4], hi[4], lo[4];
    for (int i = 0; i < 4; i++)
    {
      v[i] = w[(k * C + o) * C + chunk * 4 + i];
      hi[i] = tf32_hi(v[i]);
      lo[i] = v[i] - hi[i];
    }
    b_hi[idx] = make_float4(mode ? hi[0] : v[0], mode ? hi[1] : v[1], mode ? hi[2] : v[2], mode ? hi[3] : v[3]);
    b_lo[idx] = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of A/B must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base;

  if (tid == 0)
  {
    const uint32_t idesc = make_idesc(128, C);
    uint32_t acc = 0;
    for (int k = 0; k < NTAP; k++)
    {
      const int off = (NTAP - 1 - k) * dilation; // tap k looks back (K-1-k)*d
      for (int half = 0; half < C / 8; half++) // K = 8 input channels = 2 planes per MMA
      {
        const int col0 = t0 - off;
        const uint32_t lbo = TW * 16, sbo = 128;
        const uint64_t dah = make_desc(smem_u32(&a_hi[(2 * half) * TW + col0]), lbo, sbo);
        const uint64_t dal = make_desc(smem_u32(&a_lo[(2 * half) * TW + col0]), lbo, sbo);
        const uint64_t dbh = make_desc(smem_u32(&b_hi[(k * (C / 4) + 2 * half) * C]), C * 16, 128);
        const uint64_t dbl = make_desc(smem_u32(&b_lo[(k * (C / 4) + 2 * half) * C]), C * 16, 128);
        if (mode)
        {
          mma_tf32(tmem_d, dal, dbh, idesc, acc); // small terms first
          mma_tf32(tmem_d, dah, dbl, idesc, 1);
          mma_tf32(tmem_d, dah, dbh, idesc, 1);
        }
        else
          mma_tf32(tmem_d, dah, dbh, idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done)
      asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(&mbar)), "r"(0u));
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[16];
  const uint32_t taddr = tmem_d + ((uint32_t)(tid & ~31) << 16); // lane field: this warp's 32-lane quarter
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int o = 0; o < 16; o++)
    out[tid * C + o] = __uint_as_float(r[o]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d));
}

// fp32 FMA reference on the device (what the current kernel computes)
__global__ void fp32_kernel(const float* h, const float* w, float* out, int dilation, int t0)
{
  const int t = threadIdx.x;
  for (int o = 0; o < C; o++)
  {
    float acc = 0.f;
    for (int k = 0; k < NTAP; k++)
      for (int i = 0; i < C; i++)
        acc = fmaf(w[(k * C + o) * C + i], h[i * TW + t0 + t - (NTAP - 1 - k) * dilation], acc);
    out[t * C + o] = acc;
  }
}

int main()
{
  std::vector<float> h(C * TW), w(NTAP * C * C);
  srand(1234);
  for (auto& v : h)
    v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f; // activations in (-1, 1)
  for (auto& v : w)
    v = (rand() / (float)RAND_MAX - 0.5f) * 0.6f; // weights U(-0.3, 0.3)
  float *dh, *dw, *dout;
  CK(cudaMalloc(&dh, h.size() * 4));
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMalloc(&dout, 128 * C * 4));
  CK(cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> got(128 * C);
  for (int dilation : {1, 3, 8, 29})
  {
    const int t0 = HALO + 5; // deliberately not a multiple of 8
    std::vector<double> ref(128 * C);
    double scale = 0;
    for (int t = 0; t < 128; t++)
      for (int o = 0; o < C; o++)
      {
        double acc = 0;
        for (int k = 0; k < NTAP; k++)
          for (int i = 0; i < C; i++)
            acc += (double)w[(k * C + o) * C + i] * (double)h[i * TW + t0 + t - (NTAP - 1 - k) * dilation];
        ref[t * C + o] = acc;
        scale = fmax(scale, fabs(acc));
      }
    for (int mode = 0; mode < 3; mode++)
    {
      CK(cudaMemset(dout, 0, 128 * C * 4));
      if (mode < 2)
        probe_kernel<<<1, 128>>>(dh, dw, dout, dilation, mode, t0);
      else
        fp32_kernel<<<1, 128>>>(dh, dw, dout, dilation, t0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess)
      {
        printf("{\"test\": \"tcgen05_tap_gemm\", \"dilation\": %d, \"mode\": %d, \"error\": \"%s\"}\n", dilation, mode,
               cudaGetErrorString(e));
        return 1;
      }
      CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
      double maxabs = 0, sumsq = 0;
      for (size_t i = 0; i < got.size(); i++)
      {
        const double d = fabs((double)got[i] - ref[i]);
        maxabs = fmax(maxabs, d);
        sumsq += d * d;
      }
      const char* name = mode == 0 ? "tf32_1pass" : mode == 1 ? "tf32_3pass_split" : "fp32_fma";
      printf("{\"test\": \"tcgen05_tap_gemm\", \"dilation\": %d, \"mode\": \"%s\", \"max_abs_err\": %.3e, \"rms_err\": %.3e, "
             "\"max_abs_ref\": %.3f, \"sample\": [%.6f, %.6f], \"ref\": [%.6f, %.6f]}\n",
             dilation, name, maxabs, sqrt(sumsq / got.size()), scale, got[0], got[17 * C + 5], ref[0], ref[17 * C + 5]);
    }
  }
  return 0;
}
