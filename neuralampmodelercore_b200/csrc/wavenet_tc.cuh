// wavenet_tc.cuh -- tensor-core (tcgen05 / TMEM) variant of the fused WaveNet kernel for sm_100a.
//
// Same reference code replaced as wavenet_fused.cuh (NAM/wavenet/model.cpp:183-393,463-549,822-910,
// NAM/conv1d.cpp:163-183,666-683, NAM/dsp.cpp:436-836), same per-stream ring state -- but the two matrix
// products of a layer run on the 5th-generation tensor cores:
//
//   Z[128 x 16] = sum_taps  H_l[t - off, :] . W_tap          conv   (tcgen05.mma kind::tf32, D in TMEM)
//   D[128 x 16] = act(Z) . P                                 layer1x1
//
// with M = 128 = time steps of the CTA's tile (TMEM lane == time step), N = 16 output channels, K = 8 input
// channels per instruction.  The activation layout "planes of 4 channels, [C/4][time][4 floats]" is exactly the
// UMMA K-major no-swizzle canonical layout (core matrix = 8 consecutive time steps x 16 B), so a dilated tap is
// nothing but a shared-memory descriptor whose start address is shifted by `off` rows: no im2col, no copies --
// all taps accumulate into one TMEM tile.
//
// Precision: 1e-5 parity forbids single-pass TF32 (1.7e-3 error, profiles/r01_tc_probe_tf32_split.jsonl).
// Each product is split x = hi + lo (hi = x rounded to TF32, lo = x - hi) and evaluated as
// lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM: measured 2.4e-7 rms / 1e-6 max on O(1) outputs,
// ~3x the rounding noise of an fp32 FMA chain.  The residual stream, head accumulator, bias/mixin adds and the
// activations stay in fp32 registers, so errors do not compound through the tensor core.
//
// Threads: 256 per CTA, two per time step.  Thread (row, half) owns channels [half*C/2, (half+1)*C/2) of its
// row: warps w and w+4 read the same 32 TMEM lanes, different columns.  That halves every epilogue and doubles
// the warps available to hide the L2 / TMEM / MUFU latencies (the first version, one thread per row, sat at 29 %
// issue-active: profiles/r01c_tc_kernel_v1.json).
//
// History: this kernel's rings hold the layer inputs already split (hi planes, then lo planes), so halo columns
// and far taps move ring -> shared memory with cp.async alone.  Epilogue arithmetic runs on packed f32x2 pairs
// (the FMA pipe is idle here; issue slots are the scarce resource).
//
// Per layer:  wait for the cp.async groups (weights, halo, staged taps) -> barrier -> thread 0 issues the conv
// MMAs -> commit;   meanwhile all threads stream in the next layer's weights and persist the tail of h_l ->
// wait -> start the next layer's halo copies -> tcgen05.ld Z, +bias +mixin, activation, head += a, write a (hi/lo)
// to shared -> barrier -> 1x1 MMAs -> wait -> tcgen05.ld D, residual add in registers, write h_{l+1} (hi/lo) to the
// tile, start the next layer's far-tap copies.
#pragma once

#include "wavenet_fused.cuh"

namespace namb200
{

constexpr int kTcM = 128; // frames per tile = TMEM lanes
constexpr int kTcThreads = 256; // two threads per frame
constexpr int kTcTW = kHalo + kTcM; // tile columns per 4-channel plane

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Shared-memory matrix descriptor (K-major, SWIZZLE_NONE, version 1 = sm_100).  Low word: start address >> 4
// in bits 0-13, leading-dimension byte offset >> 4 in bits 16-29; moving the start by n float4 is `lo + n`.
// High word: stride-dimension byte offset (128 B between 8-row core matrices) >> 4, version bit 46.
__device__ __forceinline__ uint32_t tc_desc_lo(const void* p, uint32_t lbo_bytes)
{
  return ((tc_smem_u32(p) & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
constexpr uint32_t kTcDescHi = (128u >> 4) | (1u << 14);
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t accumulate)
{
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    ".reg .b64 da, db;\n\t"
    "mov.b64 da, {%1, %6};\n\t"
    "mov.b64 db, {%2, %6};\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, {%5, %5, %5, %5}, p;\n\t"
    "}\n" ::"r"(tmem_d),
    "r"(da_lo), "r"(db_lo), "r"(kTcIdesc), "r"(accumulate), "r"(0u), "r"(kTcDescHi)
    : "memory");
}
// one logical product A.B as three TF32 MMAs, smallest terms first
__device__ __forceinline__ void tc_mma3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                        uint32_t accumulate)
{
  tc_mma(tmem_d, a_lo, b_hi, accumulate);
  tc_mma(tmem_d, a_hi, b_lo, 1u);
  tc_mma(tmem_d, a_hi, b_hi, 1u);
}
__device__ __forceinline__ void tc_commit(uint64_t* mbar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(mbar))
               : "memory");
}
// One lane per warp polls the barrier (a polling warp costs issue slots that co-resident CTAs need).
__device__ __forceinline__ void tc_mbar_wait(uint64_t* mbar, uint32_t parity)
{
  if ((threadIdx.x & 31) == 0)
  {
    uint32_t done = 0;
    while (!done)
      asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(tc_smem_u32(mbar)), "r"(parity)
        : "memory");
  }
  __syncwarp();
}
__device__ __forceinline__ void tc_fence_before()
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the tensor core
}
template <int N>
__device__ __forceinline__ void tc_ld(uint32_t taddr, float (&v)[N])
{
  static_assert(N == 4 || N == 8, "tcgen05.ld width");
  uint32_t r[8];
  if constexpr (N == 8)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  else
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; i++)
    v[i] = __uint_as_float(r[i]);
}
// round to TF32, ties away from zero (what cvt.rna.tf32.f32 does, minus its Inf/NaN special-casing which
// costs three more instructions per element and cannot matter: a non-finite input stays non-finite)
__device__ __forceinline__ float tc_hi(float x)
{
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void tc_split4(const float4& v, float4& hi, float4& lo)
{
  hi = make_float4(tc_hi(v.x), tc_hi(v.y), tc_hi(v.z), tc_hi(v.w));
  lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}
__device__ __forceinline__ void tc_cp_async16(void* dst_smem, const void* src_gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc_smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}

// shared-memory carve-up shared by the kernel and its helpers
struct TcSmem
{
  float4* wbuf; // [2][wimg4] double-buffered per-layer B images
  int wimg4;
  float4* tile_hi; // [P][kTcTW]
  float4* tile_lo;
  float4* ubuf; // a_hi [P][128] | a_lo [P][128]   aliased with   stage[2 slots][hi|lo][P][128]
  float4* xch; // end of a layer array: the full fp32 residual rows [P][128] for the next array's rechannel ...
  float4* hx; // ... and the partner thread's head-rechannel partial sums (both alias ubuf)
  uint64_t* mbar_z;
  uint64_t* mbar_h;
  uint32_t tmem;
};

// start streaming layer `gl`'s B image into its buffer (all threads)
__device__ __forceinline__ void tc_prefetch_weights(const WaveNetKernelParams& p, const TcSmem& sm, int gl, uint32_t buf)
{
  const float4* src = reinterpret_cast<const float4*>(p.tc_blob + p.tc_off[gl]);
  float4* dst = sm.wbuf + buf * sm.wimg4;
  const int n4 = p.tc_floats[gl] >> 2;
  for (int i = threadIdx.x; i < n4; i += kTcThreads)
    tc_cp_async16(dst + i, src + i);
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// this thread's PH planes (NP = 2 PH pairs) of one column, split into hi / lo
template <int PH>
__device__ __forceinline__ void tc_store_pairs(float4* __restrict__ hi_base, float4* __restrict__ lo_base, int stride,
                                               int pl0, int col, const u64 (&v)[2 * PH])
{
#pragma unroll
  for (int q = 0; q < PH; q++)
  {
    ulonglong2 hi, lo;
    tc_split2(v[2 * q], hi.x, lo.x);
    tc_split2(v[2 * q + 1], hi.y, lo.y);
    *reinterpret_cast<ulonglong2*>(hi_base + (pl0 + q) * stride + col) = hi;
    *reinterpret_cast<ulonglong2*>(lo_base + (pl0 + q) * stride + col) = lo;
  }
}

// The rings of this kernel hold the layer inputs already split: hi planes [P][R], then lo planes [P][R]
// (twice the floats of the FP32 kernel's rings, at float offset 2 * ring_off of the stream's state).  History then
// moves ring -> shared memory with cp.async only: no registers, no arithmetic.
template <int C>
__device__ __forceinline__ const float4* tc_ring(const LayerDesc& L, const float* __restrict__ state)
{
  return reinterpret_cast<const float4*>(state + 2 * (size_t)L.ring_off);
}

// halo columns [t0 - halo, t0) of layer L's input: one (plane, column) item per thread, hi and lo
template <int C>
__device__ __forceinline__ void tc_halo_async(const LayerDesc& L, const TcSmem& sm, const float* __restrict__ state,
                                              uint32_t tabs0)
{
  constexpr int P = C / 4;
  const int halo = L.lookback < kHalo ? L.lookback : kHalo;
  const int col = threadIdx.x & (kHalo - 1), pl = threadIdx.x >> 6; // kHalo == 64
  if (pl < P && col >= kHalo - halo)
  {
    const int R = L.ring_mask + 1;
    const float4* src = tc_ring<C>(L, state) + pl * R + ((tabs0 - (uint32_t)kHalo + (uint32_t)col) & (uint32_t)L.ring_mask);
    tc_cp_async16(sm.tile_hi + pl * kTcTW + col, src);
    tc_cp_async16(sm.tile_lo + pl * kTcTW + col, src + P * R);
  }
}

// Taps whose window [t0 - off, t0 - off + 128) is not inside halo + tile (off > 64) are staged as their own
// A operand, at most two per layer (wavenet_pack.cpp refuses more), in tap order.  Rows before t0 come from the
// ring (cp.async); rows inside the current tile (only when 64 < off < 128) are copied from the tile once it is
// complete.
__device__ __forceinline__ void tc_staged_offsets(const LayerDesc& L, int (&off)[2])
{
  off[0] = off[1] = 0;
  int n = 0;
  for (int k = 0; k < L.kernel - 1; k++)
  {
    const int o = (L.kernel - 1 - k) * L.dilation;
    if (o > kHalo)
    {
      if (n == 0)
        off[0] = o;
      else if (n == 1)
        off[1] = o;
      n++;
    }
  }
}
// returns true when some rows still have to be copied from the tile (tc_stage_copy, after a barrier)
template <int C>
__device__ __forceinline__ bool tc_stage_async(const LayerDesc& L, const TcSmem& sm, const float* __restrict__ state,
                                               uint32_t tabs0, int row, int pl0, int (&off)[2])
{
  constexpr int PH = C / 8, P = C / 4;
  tc_staged_offsets(L, off);
  const int R = L.ring_mask + 1;
  const float4* ring = tc_ring<C>(L, state);
  bool reads_tile = false;
#pragma unroll
  for (int slot = 0; slot < 2; slot++)
  {
    const int o = off[slot];
    reads_tile |= (o != 0) && (o < kTcM);
    if (o != 0 && row - o < 0)
    {
      float4* st_hi = sm.ubuf + slot * (2 * P * kTcM);
      float4* st_lo = st_hi + P * kTcM;
#pragma unroll
      for (int q = 0; q < PH; q++)
      {
        const float4* src = ring + (pl0 + q) * R + ((tabs0 + (uint32_t)(row - o)) & (uint32_t)L.ring_mask);
        tc_cp_async16(st_hi + (pl0 + q) * kTcM + row, src);
        tc_cp_async16(st_lo + (pl0 + q) * kTcM + row, src + P * R);
      }
    }
  }
  return reads_tile;
}
template <int C>
__device__ __forceinline__ void tc_stage_copy(const int (&off)[2], const TcSmem& sm, int row, int pl0)
{
  constexpr int PH = C / 8, P = C / 4;
#pragma unroll
  for (int slot = 0; slot < 2; slot++)
  {
    const int rel = row - off[slot];
    if (off[slot] != 0 && rel >= 0)
    {
      float4* st_hi = sm.ubuf + slot * (2 * P * kTcM);
      float4* st_lo = st_hi + P * kTcM;
#pragma unroll
      for (int q = 0; q < PH; q++)
      {
        st_hi[(pl0 + q) * kTcM + row] = sm.tile_hi[(pl0 + q) * kTcTW + kHalo + rel];
        st_lo[(pl0 + q) * kTcM + row] = sm.tile_lo[(pl0 + q) * kTcTW + kHalo + rel];
      }
    }
  }
}

// One layer array on one tile.  `head` (this thread's half of the head accumulator) comes in holding the
// previous array's head output, `headout` leaves with this array's: HOUT == 1 -> the model output (valid in the
// half == 0 thread of each row), else this thread's half of the next array's head input.  For CIN > 1 the input
// rows are read from sm.xch, where the previous array left them; if HOUT > 1 this array leaves its own there.
template <int CIN, int C, int HOUT>
__device__ __forceinline__ void tc_array_forward(const WaveNetKernelParams& p, const ArrayDesc& A, const TcSmem& sm,
                                                 float* __restrict__ state, const uint32_t tabs0, const int Tv,
                                                 const int total_layers, uint32_t& par_z, uint32_t& par_h,
                                                 const float x, float (&head)[C / 2],
                                                 float (&headout)[HOUT == 1 ? 1 : HOUT / 2])
{
  constexpr int P = C / 4; // planes
  constexpr int CH = C / 2; // channels per thread
  constexpr int PH = C / 8; // planes per thread
  constexpr int KS = C / 8; // K-steps (8 input channels each) per tap
  const int tid = threadIdx.x;
  const int row = tid & (kTcM - 1), hf = tid >> 7;
  const int pl0 = hf * PH, ch0 = hf * CH;
  const uint32_t tlane = sm.tmem + ((uint32_t)(row & ~31) << 16) + (uint32_t)ch0;
  const float* __restrict__ gw = p.weights; // FFMA blob: rechannel / head weights (uniform, L1-resident)

  // ---- rechannel (Conv1x1 without bias, model.cpp:492), thread-local on this thread's output channels
  constexpr int NP = CH / 2; // packed pairs per thread
  u64 hres[NP]; // residual stream, fp32 pairs
  u64 head2[NP];
#pragma unroll
  for (int q = 0; q < NP; q++)
    head2[q] = pack2(head[2 * q], head[2 * q + 1]);
  {
    float h[CH];
#pragma unroll
    for (int o = 0; o < CH; o++)
      h[o] = 0.0f;
    if constexpr (CIN == 1)
    {
#pragma unroll
      for (int o = 0; o < CH; o++)
        h[o] = __ldg(gw + A.rech_off + ch0 + o) * x;
    }
    else
    {
#pragma unroll
      for (int pi = 0; pi < CIN / 4; pi++)
      {
        const float4 v = sm.xch[pi * kTcM + row];
        const float in4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int o = 0; o < CH; o++)
            h[o] = fmaf(__ldg(gw + A.rech_off + (4 * pi + i) * C + ch0 + o), in4[i], h[o]);
      }
    }
#pragma unroll
    for (int q = 0; q < NP; q++)
      hres[q] = pack2(h[2 * q], h[2 * q + 1]);
  }
  __syncthreads(); // xch / hx (aliases of ubuf) have been consumed by every thread; ubuf may be rewritten
  tc_store_pairs<PH>(sm.tile_hi, sm.tile_lo, kTcTW, pl0, kHalo + row, hres);
  {
    const LayerDesc& L0 = p.layers[A.layer0];
    int off0[2];
    tc_halo_async<C>(L0, sm, state, tabs0);
    if (tc_stage_async<C>(L0, sm, state, tabs0, row, pl0, off0))
    {
      __syncthreads();
      tc_stage_copy<C>(off0, sm, row, pl0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const u64 xx = tc_dup(x);

#pragma unroll 1
  for (int li = 0; li < A.n_layers; li++)
  {
    const int gl = A.layer0 + li;
    const LayerDesc& Ld = p.layers[gl];
    const bool has_next = li + 1 < A.n_layers;
    const LayerDesc& Ln = p.layers[has_next ? gl + 1 : gl];
    const int K = Ld.kernel, dil = Ld.dilation;
    const uint32_t wsel = par_z; // weight double buffer: toggles once per processed layer, like the Z barrier phase
    const float4* __restrict__ img = sm.wbuf + wsel * sm.wimg4;
    const ulonglong2* __restrict__ vec2 =
      reinterpret_cast<const ulonglong2*>(img + ((2 * K * KS + 2 * KS) * kTcTile) / 4) + pl0; // b | M | p | slopes

    asm volatile("cp.async.wait_group 0;" ::: "memory"); // this layer's B image, halo and staged taps have landed
    tc_fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- conv: all taps accumulate into Z (TMEM columns [0,16))
    if (tid == 0)
    {
      tc_fence_after();
      const uint32_t a_tile_hi = tc_desc_lo(sm.tile_hi, kTcTW * 16), a_tile_lo = tc_desc_lo(sm.tile_lo, kTcTW * 16);
      uint32_t a_st = tc_desc_lo(sm.ubuf, kTcM * 16);
      uint32_t b_hi = tc_desc_lo(img, 256);
      const uint32_t b_lo_off = (uint32_t)(K * KS * kTcTile / 4);
      uint32_t acc = 0;
      for (int k = 0; k < K; k++)
      {
        const int off = (K - 1 - k) * dil;
        uint32_t ah, al, step;
        if (off <= kHalo)
        {
          ah = a_tile_hi + (uint32_t)(kHalo - off);
          al = a_tile_lo + (uint32_t)(kHalo - off);
          step = 2 * kTcTW;
        }
        else
        {
          ah = a_st;
          al = a_st + P * kTcM;
          step = 2 * kTcM;
          a_st += 2 * P * kTcM;
        }
#pragma unroll
        for (int s = 0; s < KS; s++)
        {
          tc_mma3(sm.tmem, ah + s * step, al + s * step, b_hi, b_hi + b_lo_off, acc);
          acc = 1;
          b_hi += kTcTile / 4;
        }
      }
      tc_commit(sm.mbar_z);
    }
    // ---- while the tensor core works: stream in the next layer's weights, persist the tail of h_l (split form:
    //      this thread's own columns, read back from the tile)
    tc_prefetch_weights(p, sm, (gl + 1 == total_layers) ? 0 : gl + 1, wsel ^ 1u);
    if (row < Tv && row >= Tv - Ld.lookback)
    {
      const int R = Ld.ring_mask + 1;
      float4* __restrict__ ring = const_cast<float4*>(tc_ring<C>(Ld, state));
      const uint32_t slot = (tabs0 + (uint32_t)row) & (uint32_t)Ld.ring_mask;
#pragma unroll
      for (int q = 0; q < PH; q++)
      {
        st_ring(ring + (pl0 + q) * R + slot, sm.tile_hi[(pl0 + q) * kTcTW + kHalo + row]);
        st_ring(ring + (P + pl0 + q) * R + slot, sm.tile_lo[(pl0 + q) * kTcTW + kHalo + row]);
      }
    }
    tc_mbar_wait(sm.mbar_z, par_z);
    par_z ^= 1u;
    tc_fence_after();
    // the conv MMAs were the last readers of the halo: fetch the next layer's
    if (has_next)
      tc_halo_async<C>(Ln, sm, state, tabs0);

    // ---- epilogue 1: z = Z + b + M c ; a = act(z) ; head += a ; a -> shared (hi / lo)
    u64 a2[NP];
    {
      float a[CH];
      tc_ld<CH>(tlane, a);
#pragma unroll
      for (int q = 0; q < PH; q++)
      {
        const ulonglong2 b2 = vec2[q], m2 = vec2[4 + q];
        a2[2 * q] = tc_add2(pack2(a[4 * q], a[4 * q + 1]), tc_fma2(m2.x, xx, b2.x));
        a2[2 * q + 1] = tc_add2(pack2(a[4 * q + 2], a[4 * q + 3]), tc_fma2(m2.y, xx, b2.y));
      }
      if (Ld.act == KACT_FASTTANH)
      {
#pragma unroll
        for (int q = 0; q < NP; q++)
          a2[q] = tc_fast_tanh2(a2[q]);
      }
      else
      {
#pragma unroll
        for (int q = 0; q < NP; q++)
          unpack2(a2[q], a[2 * q], a[2 * q + 1]);
        apply_activation<CH>(a, Ld, reinterpret_cast<const float*>(vec2 + 12));
#pragma unroll
        for (int q = 0; q < NP; q++)
          a2[q] = pack2(a[2 * q], a[2 * q + 1]);
      }
    }
#pragma unroll
    for (int q = 0; q < NP; q++)
      head2[q] = tc_add2(head2[q], a2[q]); // model.cpp:530
    tc_store_pairs<PH>(sm.ubuf, sm.ubuf + P * kTcM, kTcM, pl0, row, a2);
    tc_fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- layer1x1: D = a . P (TMEM columns [16,32))
    if (tid == 0)
    {
      tc_fence_after();
      const uint32_t a_hi = tc_desc_lo(sm.ubuf, kTcM * 16);
      const uint32_t b_hi = tc_desc_lo(img + (2 * K * KS * kTcTile) / 4, 256);
#pragma unroll
      for (int s = 0; s < KS; s++)
        tc_mma3(sm.tmem + 16, a_hi + s * 2 * kTcM, a_hi + P * kTcM + s * 2 * kTcM, b_hi + s * (kTcTile / 4),
                b_hi + (KS + s) * (kTcTile / 4), s > 0 ? 1u : 0u);
      tc_commit(sm.mbar_h);
    }
    tc_mbar_wait(sm.mbar_h, par_h);
    par_h ^= 1u;
    tc_fence_after();

    // ---- epilogue 2: h_{l+1} = h_l + p + D   (model.cpp:243,376), fp32 in registers
    {
      float d[CH];
      tc_ld<CH>(tlane + 16, d);
#pragma unroll
      for (int q = 0; q < PH; q++)
      {
        const ulonglong2 p2 = vec2[8 + q];
        hres[2 * q] = tc_add2(hres[2 * q], tc_add2(p2.x, pack2(d[4 * q], d[4 * q + 1])));
        hres[2 * q + 1] = tc_add2(hres[2 * q + 1], tc_add2(p2.y, pack2(d[4 * q + 2], d[4 * q + 3])));
      }
    }
    if (has_next)
    {
      tc_store_pairs<PH>(sm.tile_hi, sm.tile_lo, kTcTW, pl0, kHalo + row, hres);
      // the 1x1 MMAs were the last readers of ubuf: the next layer's far taps may land there now
      int offn[2];
      if (tc_stage_async<C>(Ln, sm, state, tabs0, row, pl0, offn))
      {
        __syncthreads(); // the rows copied from the tile were written by other threads just now
        tc_stage_copy<C>(offn, sm, row, pl0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll
  for (int q = 0; q < NP; q++)
    unpack2(head2[q], head[2 * q], head[2 * q + 1]);
  float hres_f[CH];
#pragma unroll
  for (int q = 0; q < NP; q++)
    unpack2(hres[q], hres_f[2 * q], hres_f[2 * q + 1]);

  // ---- head rechannel (kernel size 1; model.cpp:548): partial sums over this thread's channels, the partner
  //      thread of the row supplies the other half through shared memory
  const float* __restrict__ wh = gw + A.head_off;
  if constexpr (HOUT == 1)
  {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; i++)
      s = fmaf(__ldg(wh + ch0 + i), head[i], s);
    float* hx = reinterpret_cast<float*>(sm.hx);
    if (hf == 1)
      hx[row] = s;
    __syncthreads();
    headout[0] = (hf == 0) ? (s + hx[row]) + __ldg(wh + C) : 0.0f;
  }
  else
  {
    constexpr int HH = HOUT / 2, PHN = HOUT / 8;
    float part[HOUT];
#pragma unroll
    for (int ho = 0; ho < HOUT; ho++)
      part[ho] = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; i++)
#pragma unroll
      for (int ho = 0; ho < HOUT; ho++)
        part[ho] = fmaf(__ldg(wh + (ch0 + i) * HOUT + ho), head[i], part[ho]);
    // leave the partner's channels in hx, this array's residual rows in xch
#pragma unroll
    for (int q = 0; q < PHN; q++)
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hf == 0)
        v = make_float4(part[HH + 4 * q], part[HH + 4 * q + 1], part[HH + 4 * q + 2], part[HH + 4 * q + 3]);
      else
        v = make_float4(part[4 * q], part[4 * q + 1], part[4 * q + 2], part[4 * q + 3]);
      sm.hx[((1 - hf) * PHN + q) * kTcM + row] = v;
    }
#pragma unroll
    for (int q = 0; q < PH; q++)
      sm.xch[(pl0 + q) * kTcM + row] =
        make_float4(hres_f[4 * q], hres_f[4 * q + 1], hres_f[4 * q + 2], hres_f[4 * q + 3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PHN; q++)
    {
      const float4 o4 = sm.hx[(hf * PHN + q) * kTcM + row];
      const float oth[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
      {
        const float own = (hf == 0) ? part[4 * q + i] : part[HH + 4 * q + i];
        headout[4 * q + i] = (own + oth[i]) + __ldg(wh + C * HOUT + hf * HH + 4 * q + i);
      }
    }
  }
}

// One persistent CTA (256 threads, one 128-frame tile at a time) per stream slot.
template <int C0, int C1>
__global__ void __launch_bounds__(kTcThreads, 3) wavenet_tc_kernel(const __grid_constant__ WaveNetKernelParams p,
                                                                   const int wimg4, const int total_layers)
{
  constexpr int CMAX = (C0 > C1) ? C0 : C1;
  constexpr int PM = CMAX / 4;
  extern __shared__ float4 smem4[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base;
  TcSmem sm;
  sm.wbuf = smem4;
  sm.wimg4 = wimg4;
  sm.tile_hi = sm.wbuf + 2 * wimg4;
  sm.tile_lo = sm.tile_hi + PM * kTcTW;
  sm.ubuf = sm.tile_lo + PM * kTcTW; // 2 slots x (hi|lo) x PM planes x 128
  sm.xch = sm.ubuf;
  sm.hx = sm.ubuf + PM * kTcM;
  sm.mbar_z = &mbar[0];
  sm.mbar_h = &mbar[1];
  const int tid = threadIdx.x;
  const int row = tid & (kTcM - 1);

  if (tid == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(tc_smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  sm.tmem = tmem_base;
  uint32_t par_z = 0, par_h = 0;
  tc_prefetch_weights(p, sm, 0, 0u);

  for (int stream = blockIdx.x; stream < p.batch; stream += gridDim.x)
  {
    float* __restrict__ state = p.state + (size_t)stream * p.state_stride;
    const float* __restrict__ xin = p.in + (size_t)stream * p.in_stride;
    float* __restrict__ yout = p.out + (size_t)stream * p.out_stride;
    for (int t0 = 0; t0 < p.n_frames; t0 += kTcM)
    {
      const int Tv = min(kTcM, p.n_frames - t0);
      const uint32_t tabs0 = p.t_base + (uint32_t)t0;
      // no condition_dsp: condition == input (model.cpp:781)
      const float x = (row < Tv) ? __ldg(xin + t0 + row) : 0.0f;
      float y;
      if constexpr (C1 == 0)
      {
        float head0[C0 / 2], ho0[1];
#pragma unroll
        for (int o = 0; o < C0 / 2; o++)
          head0[o] = 0.0f;
        tc_array_forward<1, C0, 1>(p, p.arrays[0], sm, state, tabs0, Tv, total_layers, par_z, par_h, x, head0, ho0);
        y = ho0[0];
      }
      else
      {
        float ho0[C1 / 2], ho1[1];
        {
          float head0[C0 / 2];
#pragma unroll
          for (int o = 0; o < C0 / 2; o++)
            head0[o] = 0.0f;
          tc_array_forward<1, C0, C1>(p, p.arrays[0], sm, state, tabs0, Tv, total_layers, par_z, par_h, x, head0, ho0);
        }
        tc_array_forward<C0, C1, 1>(p, p.arrays[1], sm, state, tabs0, Tv, total_layers, par_z, par_h, x, ho0, ho1);
        y = ho1[0];
      }
      if (tid < Tv) // tid < 128: the half-0 thread of the row holds the output
        yout[t0 + tid] = p.head_scale * y; // model.cpp:888-897
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(sm.tmem));
}

} // namespace namb200
