#include "wavenet_pack.h"

#include <algorithm>
#include <cstdint>
#include <cstring>

namespace namb200
{

namespace
{
void pack_tc(const ModelSpec& ms, WaveNetPlan& plan);

int pad_channels(int c)
{
  if (c <= 4)
    return 4;
  if (c <= 8)
    return 8;
  if (c <= 16)
    return 16;
  return -1;
}

int next_pow2(long v)
{
  int r = 1;
  while (r < v)
    r <<= 1;
  return r;
}

size_t align4(size_t n)
{
  return (n + 3) & ~(size_t)3;
}

// Round to TF32 (10 explicit mantissa bits), nearest, ties away from zero -- the host twin of
// cvt.rna.tf32.f32, used to split a weight into hi + lo for the 3xTF32 tensor-core product.
float tf32_rna(float x)
{
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u)
    return x; // inf / nan
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

// Write one B tile: [2 chunks of 4 input channels][16 output channels][4] for K-step `ks` of matrix
// W (out x in, row-major, `cr` real channels), hi and lo parts.
void write_b_tile(float* hi, float* lo, const float* w_out_in, int cr_out, int cr_in, int ks)
{
  for (int c = 0; c < 2; c++)
    for (int n = 0; n < 16; n++)
      for (int i = 0; i < 4; i++)
      {
        const int in = 8 * ks + 4 * c + i;
        float v = 0.0f;
        if (n < cr_out && in < cr_in)
          v = w_out_in[(size_t)n * cr_in + in];
        const float h = tf32_rna(v);
        hi[(c * 16 + n) * 4 + i] = h;
        lo[(c * 16 + n) * 4 + i] = v - h;
      }
}

// Tensor-core images (layout: wavenet_desc.h).  Eligible when every array is padded to 8 or 16 channels
// and no layer needs more than 2 staged taps (kernel size 3 with look-back beyond the 64-column halo).
void pack_tc(const ModelSpec& ms, WaveNetPlan& plan)
{
  const WaveNetSpec& wn = ms.wavenet;
  plan.tc_eligible = false;
  for (int a = 0; a < plan.n_arrays; a++)
    if (plan.cp[a] < 8)
    {
      plan.tc_why_not = "fewer than 5 channels";
      return;
    }
  for (const ArraySpec& A : wn.arrays)
    if (A.head_kernel != 1)
    {
      plan.tc_why_not = "head kernel size " + std::to_string(A.head_kernel);
      return;
    }
  size_t li = 0;
  for (size_t a = 0; a < wn.arrays.size(); a++)
  {
    const ArraySpec& A = wn.arrays[a];
    const int cp = plan.cp[a], cr = A.channels;
    for (const LayerSpec& L : A.layers)
    {
      const int K = L.conv.kernel, ks = cp / 8;
      int staged = 0;
      for (int k = 0; k < K - 1; k++)
        if ((K - 1 - k) * L.conv.dilation > kHalo)
          staged++;
      plan.tc_max_staged_taps = std::max(plan.tc_max_staged_taps, staged);
      const int n = tc_image_floats(K, cp);
      const size_t off = plan.tc_blob.size();
      plan.tc_blob.resize(off + n, 0.0f);
      float* img = plan.tc_blob.data() + off;
      float* conv_hi = img;
      float* conv_lo = img + (size_t)K * ks * kTcTile;
      float* p_hi = img + (size_t)2 * K * ks * kTcTile;
      float* p_lo = p_hi + (size_t)ks * kTcTile;
      float* vec = p_lo + (size_t)ks * kTcTile;
      for (int k = 0; k < K; k++)
        for (int s = 0; s < ks; s++)
          write_b_tile(conv_hi + ((size_t)k * ks + s) * kTcTile, conv_lo + ((size_t)k * ks + s) * kTcTile,
                       L.conv.w.data() + (size_t)k * cr * cr, cr, cr, s);
      for (int s = 0; s < ks; s++)
        write_b_tile(p_hi + (size_t)s * kTcTile, p_lo + (size_t)s * kTcTile, L.l1x1.w.data(), cr, cr, s);
      for (int o = 0; o < cr; o++)
      {
        vec[o] = L.conv.b[o];
        vec[16 + o] = L.mixin.w[o];
        vec[32 + o] = L.l1x1.b[o];
        if (L.act.type == ActType::PReLU)
          vec[48 + o] = L.act.slopes.size() == 1 ? L.act.slopes[0] : L.act.slopes[o];
      }
      plan.tc_off.push_back((int)off);
      plan.tc_floats.push_back(n);
      plan.tc_max_image_floats = std::max(plan.tc_max_image_floats, n);
      li++;
    }
  }
  if (plan.tc_max_staged_taps > 2)
  {
    plan.tc_why_not = "more than 2 taps per layer reach beyond the halo (kernel size > 3 with long dilation)";
    return;
  }
  plan.tc_eligible = true;
}
} // namespace

WaveNetPlan plan_wavenet(const ModelSpec& ms)
{
  WaveNetPlan plan;
  const WaveNetSpec& wn = ms.wavenet;
  auto no = [&plan](const std::string& why) {
    plan.eligible = false;
    plan.why_not = why;
    return plan;
  };
  if (ms.arch != Arch::WaveNet)
    return no("not a WaveNet");
  if (wn.in_channels != 1 || ms.out_channels != 1)
    return no("fused kernel is mono in / mono out (in_channels " + std::to_string(wn.in_channels) + ", out_channels "
              + std::to_string(ms.out_channels) + ")");
  if (wn.condition_dsp)
    return no("condition_dsp sub-model");
  if (wn.with_head)
    return no("post-stack head");
  if (wn.arrays.empty() || wn.arrays.size() > 2)
    return no(std::to_string(wn.arrays.size()) + " layer arrays (fused kernel handles 1 or 2)");
  plan.n_arrays = (int)wn.arrays.size();
  size_t total_layers = 0;
  for (size_t a = 0; a < wn.arrays.size(); a++)
  {
    const ArraySpec& A = wn.arrays[a];
    const std::string where = "layer array " + std::to_string(a) + ": ";
    if (A.condition_size != 1)
      return no(where + "condition_size " + std::to_string(A.condition_size));
    if (A.bottleneck != A.channels)
      return no(where + "bottleneck != channels");
    if (pad_channels(A.channels) < 0)
      return no(where + std::to_string(A.channels) + " channels (max 16)");
    if (A.groups_input != 1 || A.groups_input_mixin != 1 || A.l1x1_groups != 1)
      return no(where + "grouped convolutions");
    if (!A.l1x1_active)
      return no(where + "layer1x1 inactive");
    if (A.h1x1_active)
      return no(where + "head1x1");
    if ((A.head_kernel - 1) * A.head_dilation > kHalo)
      return no(where + "head convolution looks back " + std::to_string((A.head_kernel - 1) * A.head_dilation)
                + " frames (max " + std::to_string(kHalo) + ")");
    const int expect_in = (a == 0) ? 1 : wn.arrays[a - 1].channels;
    if (A.input_size != expect_in)
      return no(where + "input_size " + std::to_string(A.input_size) + " (expected " + std::to_string(expect_in) + ")");
    for (const LayerSpec& L : A.layers)
    {
      if (L.gating != Gating::None)
        return no(where + "gated / blended activation");
      for (int f = 0; f < F_COUNT; f++)
        if (L.film[f].active)
          return no(where + "FiLM");
      if (L.act.type == ActType::PReLU && L.act.slopes.size() != 1 && (int)L.act.slopes.size() != A.channels)
        return no(where + "PReLU slope count");
    }
    total_layers += A.layers.size();
    plan.cp[a] = pad_channels(A.channels);
    plan.creal[a] = A.channels;
  }
  if (wn.arrays.back().head_size != 1)
    return no("last head_size != 1");
  if (total_layers > (size_t)kMaxLayers)
    return no("too many layers");

  // ---- pack ----
  std::vector<float>& blob = plan.blob;
  long ring_off = 0;
  double macs = 0.0;
  for (size_t a = 0; a < wn.arrays.size(); a++)
  {
    const ArraySpec& A = wn.arrays[a];
    const int C = plan.cp[a], Cr = A.channels;
    const int CIN = (a == 0) ? 1 : plan.cp[a - 1];
    const int CINr = A.input_size;
    const bool is_last = (a + 1 == wn.arrays.size());
    const int HOUT = is_last ? 1 : plan.cp[a + 1];
    const int HOUTr = A.head_size;
    ArrayDesc ad;
    ad.layer0 = (int)plan.layers.size();
    ad.n_layers = (int)A.layers.size();
    // rechannel [CIN][C]
    ad.rech_off = (int)blob.size();
    blob.resize(blob.size() + (size_t)CIN * C, 0.0f);
    for (int i = 0; i < CINr; i++)
      for (int o = 0; o < Cr; o++)
        blob[ad.rech_off + (size_t)i * C + o] = A.rechannel.w[(size_t)o * CINr + i];
    macs += (double)CINr * Cr;
    for (const LayerSpec& L : A.layers)
    {
      LayerDesc ld{};
      const int K = L.conv.kernel;
      ld.w_off = (int)blob.size();
      ld.kernel = K;
      ld.dilation = L.conv.dilation;
      ld.lookback = (int)L.conv.lookback();
      if (ld.lookback > plan.max_lookback)
        plan.max_lookback = ld.lookback;
      const int R = next_pow2(ld.lookback > 0 ? ld.lookback : 1);
      ld.ring_mask = R - 1;
      ld.ring_off = (int)ring_off;
      ring_off += (long)C * R;
      ld.act = (int)L.act.type;
      ld.ap0 = ld.ap1 = ld.ap2 = ld.ap3 = 0.0f;
      if (L.act.type == ActType::LeakyReLU)
        ld.ap0 = L.act.slope;
      else if (L.act.type == ActType::LeakyHardtanh)
      {
        ld.ap0 = L.act.min_val;
        ld.ap1 = L.act.max_val;
        ld.ap2 = L.act.min_slope;
        ld.ap3 = L.act.max_slope;
      }
      // conv [K][C(in)][C(out)]
      const size_t conv_off = blob.size();
      blob.resize(blob.size() + (size_t)K * C * C, 0.0f);
      for (int k = 0; k < K; k++)
        for (int i = 0; i < Cr; i++)
          for (int o = 0; o < Cr; o++)
            blob[conv_off + ((size_t)k * C + i) * C + o] = L.conv.w[((size_t)k * Cr + o) * Cr + i];
      // bias [C]
      const size_t b_off = blob.size();
      blob.resize(blob.size() + C, 0.0f);
      for (int o = 0; o < Cr; o++)
        blob[b_off + o] = L.conv.b[o];
      // mixin [C]  (condition_size == 1)
      const size_t m_off = blob.size();
      blob.resize(blob.size() + C, 0.0f);
      for (int o = 0; o < Cr; o++)
        blob[m_off + o] = L.mixin.w[o];
      // layer1x1 P [C(in)][C(out)], bias [C]
      const size_t p_off = blob.size();
      blob.resize(blob.size() + (size_t)C * C, 0.0f);
      for (int i = 0; i < Cr; i++)
        for (int o = 0; o < Cr; o++)
          blob[p_off + (size_t)i * C + o] = L.l1x1.w[(size_t)o * Cr + i];
      const size_t pb_off = blob.size();
      blob.resize(blob.size() + C, 0.0f);
      for (int o = 0; o < Cr; o++)
        blob[pb_off + o] = L.l1x1.b[o];
      // PReLU slopes [C]
      const size_t s_off = blob.size();
      blob.resize(blob.size() + C, 0.0f);
      if (L.act.type == ActType::PReLU)
        for (int o = 0; o < Cr; o++)
          blob[s_off + o] = L.act.slopes.size() == 1 ? L.act.slopes[0] : L.act.slopes[o];
      macs += (double)K * Cr * Cr + Cr + (double)Cr * Cr;
      plan.layers.push_back(ld);
    }
    // head rechannel [HK][C(in)][HOUT] + bias[HOUT], padded to a multiple of 4 floats
    const int HK = A.head_kernel;
    ad.head_kernel = HK;
    ad.head_dilation = A.head_dilation;
    ad.head_off = (int)blob.size();
    blob.resize(blob.size() + align4((size_t)HK * C * HOUT + HOUT), 0.0f);
    for (int k = 0; k < HK; k++)
      for (int i = 0; i < Cr; i++)
        for (int o = 0; o < HOUTr; o++)
          blob[ad.head_off + ((size_t)k * C + i) * HOUT + o] = A.head_rechannel.w[((size_t)k * HOUTr + o) * Cr + i];
    if (A.head_rechannel.bias)
      for (int o = 0; o < HOUTr; o++)
        blob[ad.head_off + (size_t)HK * C * HOUT + o] = A.head_rechannel.b[o];
    macs += (double)HK * Cr * HOUTr;
    ad.head_ring_off = 0;
    ad.head_ring_mask = 0;
    if (HK > 1)
    {
      const int hl = (HK - 1) * A.head_dilation;
      const int R = next_pow2(hl);
      ad.head_ring_off = (int)ring_off;
      ad.head_ring_mask = R - 1;
      ring_off += (long)C * R;
      if (hl > plan.max_lookback)
        plan.max_lookback = hl;
    }
    plan.arrays.push_back(ad);
  }
  blob.resize(align4(blob.size()), 0.0f);
  plan.state_floats = (ring_off + 31) & ~31L;
  plan.head_scale = wn.head_scale;
  plan.macs_per_frame = macs;
  plan.eligible = true;
  pack_tc(ms, plan);
  return plan;
}

} // namespace namb200
