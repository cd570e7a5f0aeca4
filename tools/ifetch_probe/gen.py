"""Generates ifetch_probe.cu: does instruction fetch of straight-line code beyond the 32 KB L1.5 scale with the number of
independent code streams on one SM?  16 segments of NI FFMA-immediate instructions each (64 KB of code per segment at
NI = 4096); warp w of a single CTA runs segment (w % segs); cycles per instruction for several (warps, segs)."""
import sys
NI = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SEGS = 16
out = ["#include <cstdio>", "#include <cuda_runtime.h>"]
for s in range(SEGS):
    out.append(f"__device__ __noinline__ float seg{s}(float a0, float b) {{")
    out.append("  float a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;")
    for i in range(NI):
        c = 0.5 + ((s * NI + i) * 2654435761 % 1000003) / 2000006.0
        out.append(f"  a{i % 8} = fmaf(a{i % 8}, {c:.9f}f, b);")
    out.append("  return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));")
    out.append("}")
out.append("""
__global__ void probe(float* sink, long long* cycles, int segs, int reps, float b)
{
  const int w = threadIdx.x >> 5;
  const int s = w % segs;
  float a = (float)threadIdx.x * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++)
  {
    switch (s)
    {
""")
for s in range(SEGS):
    out.append(f"      case {s}: a = seg{s}(a, b); break;")
out.append("""    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0)
    cycles[blockIdx.x * 32 + w] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

int main()
{
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * sizeof(float));
  cudaMallocManaged(&cyc, 148 * 32 * sizeof(long long));
  const int NI = %d, reps = 20;
  const int cfg[][3] = {{1, 1, 1}, {2, 1, 1}, {4, 1, 1}, {8, 1, 1}, {2, 2, 1}, {4, 4, 1}, {8, 4, 1}, {8, 8, 1}, {16, 8, 1}, {16, 16, 1}, {32, 16, 1},
                        {4, 4, 148}, {8, 8, 148}, {16, 16, 148}};
  for (auto& c : cfg)
  {
    for (int it = 0; it < 2; it++)
    {
      probe<<<c[2], 32 * c[0]>>>(sink, cyc, c[1], reps, 0.25f);
      cudaDeviceSynchronize();
    }
    long long mx = 0;
    for (int w = 0; w < c[0]; w++)
      mx = cyc[w] > mx ? cyc[w] : mx;
    const double per = (double)mx / ((double)NI * reps);
    printf("{\\"warps\\": %%d, \\"segments\\": %%d, \\"ctas\\": %%d, \\"cycles_per_instruction_per_warp\\": %%.3f, \\"code_bytes_per_cycle_sm\\": %%.2f}\\n", c[0], c[1], c[2],
           per, 16.0 * c[1] / per * ((double)c[0] / c[1] >= 1 ? 1 : 1));
  }
  printf("%%s\\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
""" % NI)
open("tools/ifetch_probe/ifetch_probe.cu", "w").write("\n".join(out))
