/*
 * nam_oracle.c -- scalar fp32 CPU restatement of the NeuralAmpModelerCore inference path.
 *
 * TEST INFRASTRUCTURE ONLY (see nam_oracle.h).  Written from the behaviour of the reference,
 * not copied from it: plain C99 loops over column-major (channels x frames) blocks, one
 * function per reference module, each citing the file:line it restates (paths relative to
 * the reference repository root).
 *
 * Summation order: every matrix product is one running accumulator per output channel,
 * "for each tap k (oldest first): for each input channel i ascending: acc += w*x", starting
 * from zero, bias added last -- the order of NAM/conv1d.cpp:676-682,769 with the innermost
 * Eigen GEMM order (unspecified upstream) fixed to ascending i.  Build the checker with -ffp-contract=off so results do not depend
 * on the host's FMA support.
 */
#include "nam_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

static _Thread_local char g_err[512];

static void set_err(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* nam_oracle_last_error(void)
{
  return g_err;
}

static void* xcalloc(size_t n, size_t sz)
{
  void* p = calloc(n ? n : 1, sz ? sz : 1);
  if (!p)
  {
    fprintf(stderr, "nam_oracle: out of memory\n");
    abort();
  }
  return p;
}

/* ------------------------------------------------------------------------------------------
 * Cursors over the three flat input arrays
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  const int32_t* cfg;
  int n_cfg, i_cfg;
  const float* fp;
  int n_fp, i_fp;
  const float* w;
  int n_w, i_w;
  int failed;
} cursor_t;

static int32_t take_i(cursor_t* c)
{
  if (c->i_cfg >= c->n_cfg)
  {
    c->failed = 1;
    return 0;
  }
  return c->cfg[c->i_cfg++];
}
static float take_f(cursor_t* c)
{
  if (c->i_fp >= c->n_fp)
  {
    c->failed = 1;
    return 0.0f;
  }
  return c->fp[c->i_fp++];
}
/* The weight stream running dry is the reference's "model expects more" error
 * (NAM/wavenet/model.cpp:671-682). */
static float take_w(cursor_t* c)
{
  if (c->i_w >= c->n_w)
  {
    c->failed = 2;
    return 0.0f;
  }
  return c->w[c->i_w++];
}

/* ------------------------------------------------------------------------------------------
 * Activations  (NAM/activations.h:59-133 scalar functions, :182-369 classes)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int type;
  float slope; /* LeakyReLU */
  int n_slopes; /* PReLU */
  float* slopes;
  float min_val, max_val, min_slope, max_slope; /* LeakyHardtanh */
} act_t;

/* NAM/activations.h:91-98 */
static inline float fast_tanh_f(const float x)
{
  const float ax = fabsf(x);
  const float x2 = x * x;
  return (x * (2.45550750702956f + 2.45550750702956f * ax + (0.893229853513558f + 0.821226666969744f * ax) * x2)
          / (2.44506634652299f + (2.44506634652299f + x2) * fabsf(x + 0.814642734961073f * x * ax)));
}
/* NAM/activations.h:100-103 */
static inline float fast_sigmoid_f(const float x)
{
  return 0.5f * (fast_tanh_f(x * 0.5f) + 1.0f);
}
/* NAM/activations.h:64-67 */
static inline float sigmoid_f(const float x)
{
  return 1.0f / (1.0f + expf(-x));
}
static inline float leaky_relu_f(const float x, const float s)
{
  return x > 0.0f ? x : s * x;
}

static void act_init(act_t* a, int type)
{
  memset(a, 0, sizeof(*a));
  a->type = type;
  a->slope = 0.01f;
  a->min_val = -1.0f;
  a->max_val = 1.0f;
  a->min_slope = 0.01f;
  a->max_slope = 0.01f;
}

/* cfg: type, n_params ; fparams: the parameters.
 * Mirrors Activation::get_activation(const ActivationConfig&) NAM/activations.cpp:132-166. */
static void act_parse(act_t* a, cursor_t* c)
{
  const int type = take_i(c);
  const int np = take_i(c);
  act_init(a, type);
  if (type == NAM_ACT_LEAKYRELU && np >= 1)
  {
    a->slope = take_f(c);
    for (int i = 1; i < np; i++)
      (void)take_f(c);
  }
  else if (type == NAM_ACT_PRELU)
  {
    a->n_slopes = np > 0 ? np : 1;
    a->slopes = (float*)xcalloc((size_t)a->n_slopes, sizeof(float));
    if (np > 0)
      for (int i = 0; i < np; i++)
        a->slopes[i] = take_f(c);
    else
      a->slopes[0] = 0.01f;
  }
  else if (type == NAM_ACT_LEAKYHARDTANH && np >= 4)
  {
    a->min_val = take_f(c);
    a->max_val = take_f(c);
    a->min_slope = take_f(c);
    a->max_slope = take_f(c);
    for (int i = 4; i < np; i++)
      (void)take_f(c);
  }
  else
  {
    for (int i = 0; i < np; i++)
      (void)take_f(c);
  }
}

static void act_free(act_t* a)
{
  free(a->slopes);
  a->slopes = NULL;
}

static void act_clone(act_t* dst, const act_t* src)
{
  *dst = *src;
  if (src->slopes)
  {
    dst->slopes = (float*)xcalloc((size_t)src->n_slopes, sizeof(float));
    memcpy(dst->slopes, src->slopes, sizeof(float) * (size_t)src->n_slopes);
  }
}

/* Apply to a contiguous column-major (channels x frames) block of `size` floats.
 * PReLU picks the slope by (pos % n_slopes), NAM/activations.h:296-300. `fast_tanh` swaps
 * Tanh for the rational approximation (NAM/activations.cpp:168-177). */
static void act_apply(const act_t* a, float* d, long size, int fast_tanh)
{
  switch (a->type)
  {
    case NAM_ACT_TANH:
      if (fast_tanh)
        for (long p = 0; p < size; p++)
          d[p] = fast_tanh_f(d[p]);
      else
        for (long p = 0; p < size; p++)
          d[p] = tanhf(d[p]);
      break;
    case NAM_ACT_FASTTANH:
      for (long p = 0; p < size; p++)
        d[p] = fast_tanh_f(d[p]);
      break;
    case NAM_ACT_HARDTANH:
      for (long p = 0; p < size; p++)
      {
        const float t = d[p] < -1.0f ? -1.0f : d[p];
        d[p] = t > 1.0f ? 1.0f : t;
      }
      break;
    case NAM_ACT_RELU:
      for (long p = 0; p < size; p++)
        d[p] = d[p] > 0.0f ? d[p] : 0.0f;
      break;
    case NAM_ACT_LEAKYRELU:
      for (long p = 0; p < size; p++)
        d[p] = leaky_relu_f(d[p], a->slope);
      break;
    case NAM_ACT_PRELU:
      for (long p = 0; p < size; p++)
        d[p] = leaky_relu_f(d[p], a->slopes[p % a->n_slopes]);
      break;
    case NAM_ACT_SIGMOID:
      for (long p = 0; p < size; p++)
        d[p] = sigmoid_f(d[p]);
      break;
    case NAM_ACT_SILU:
      for (long p = 0; p < size; p++)
        d[p] = d[p] * sigmoid_f(d[p]);
      break;
    case NAM_ACT_HARDSWISH:
      for (long p = 0; p < size; p++)
      {
        const float x = d[p];
        const float t = x + 3.0f;
        const float cl = t < 0.0f ? 0.0f : (t > 6.0f ? 6.0f : t);
        d[p] = x * cl * (1.0f / 6.0f);
      }
      break;
    case NAM_ACT_LEAKYHARDTANH:
      for (long p = 0; p < size; p++)
      {
        const float x = d[p];
        if (x < a->min_val)
          d[p] = (x - a->min_val) * a->min_slope + a->min_val;
        else if (x > a->max_val)
          d[p] = (x - a->max_val) * a->max_slope + a->max_val;
      }
      break;
    case NAM_ACT_SOFTSIGN:
      for (long p = 0; p < size; p++)
        d[p] = d[p] / (1.0f + fabsf(d[p]));
      break;
    case NAM_ACT_IDENTITY:
    default: break;
  }
}

/* ------------------------------------------------------------------------------------------
 * Conv1x1  (NAM/dsp.cpp:304-398 construction/weights, :436-836 process_)
 * Dense (out x in) matrix, block diagonal when grouped; weights per group row-major
 * (out_per_group, in_per_group); optional bias.
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int in_ch, out_ch, groups, has_bias;
  float* w; /* [o*in_ch + i] */
  float* wt; /* transposed copy [i*out_ch + o]: lets the compiler vectorise over output channels */
  float* b;
} conv1x1_t;

static int conv1x1_init(conv1x1_t* m, int in_ch, int out_ch, int bias, int groups)
{
  memset(m, 0, sizeof(*m));
  if (groups <= 0 || in_ch % groups != 0 || out_ch % groups != 0)
  {
    set_err("Conv1x1: channels (%d -> %d) must be divisible by groups (%d)", in_ch, out_ch, groups);
    return -1;
  }
  m->in_ch = in_ch;
  m->out_ch = out_ch;
  m->groups = groups;
  m->has_bias = bias;
  m->w = (float*)xcalloc((size_t)in_ch * out_ch, sizeof(float));
  m->wt = (float*)xcalloc((size_t)in_ch * out_ch, sizeof(float));
  m->b = (float*)xcalloc((size_t)out_ch, sizeof(float));
  return 0;
}

static void conv1x1_set_weights(conv1x1_t* m, cursor_t* c)
{
  const int opg = m->out_ch / m->groups, ipg = m->in_ch / m->groups;
  /* depthwise (groups == in == out) stores one weight per channel: the same loop covers it */
  for (int g = 0; g < m->groups; g++)
    for (int i = 0; i < opg; i++)
      for (int j = 0; j < ipg; j++)
        m->w[(size_t)(g * opg + i) * m->in_ch + (g * ipg + j)] = take_w(c);
  if (m->has_bias)
    for (int i = 0; i < m->out_ch; i++)
      m->b[i] = take_w(c);
  for (int o = 0; o < m->out_ch; o++)
    for (int i = 0; i < m->in_ch; i++)
      m->wt[(size_t)i * m->out_ch + o] = m->w[(size_t)o * m->in_ch + i];
}

static void conv1x1_free(conv1x1_t* m)
{
  free(m->w);
  free(m->wt);
  free(m->b);
  m->w = m->wt = m->b = NULL;
}

static void conv1x1_clone(conv1x1_t* d, const conv1x1_t* s)
{
  *d = *s;
  d->w = (float*)xcalloc((size_t)s->in_ch * s->out_ch, sizeof(float));
  d->wt = (float*)xcalloc((size_t)s->in_ch * s->out_ch, sizeof(float));
  d->b = (float*)xcalloc((size_t)s->out_ch, sizeof(float));
  memcpy(d->w, s->w, sizeof(float) * (size_t)s->in_ch * s->out_ch);
  memcpy(d->wt, s->wt, sizeof(float) * (size_t)s->in_ch * s->out_ch);
  memcpy(d->b, s->b, sizeof(float) * (size_t)s->out_ch);
}

/* acc[0:oc] += W^T[i][0:oc] * x[i] for i in [0,ic).  Output channels are the vector lanes (GCC
 * vector extensions), so each lane performs exactly the scalar recurrence acc += w*x in ascending i:
 * same arithmetic and order as the generic loop, just 4/8 output channels per instruction. */
typedef float v4f __attribute__((vector_size(16), aligned(4)));
typedef float v8f __attribute__((vector_size(32), aligned(4)));

#define NAM_DEFINE_AXPY8(OC)                                                                                         \
  static inline void axpy_rows_##OC(float* restrict acc, const float* restrict wt, const float* restrict x, int ic) \
  {                                                                                                                  \
    v8f a[OC / 8];                                                                                                   \
    for (int j = 0; j < OC / 8; j++)                                                                                 \
      a[j] = *(const v8f*)(acc + 8 * j);                                                                             \
    for (int i = 0; i < ic; i++)                                                                                     \
    {                                                                                                                \
      const float xi = x[i];                                                                                         \
      const v8f xv = {xi, xi, xi, xi, xi, xi, xi, xi};                                                               \
      const float* wc = wt + (size_t)i * OC;                                                                         \
      for (int j = 0; j < OC / 8; j++)                                                                               \
        a[j] += *(const v8f*)(wc + 8 * j) * xv;                                                                      \
    }                                                                                                                \
    for (int j = 0; j < OC / 8; j++)                                                                                 \
      *(v8f*)(acc + 8 * j) = a[j];                                                                                   \
  }
NAM_DEFINE_AXPY8(8)
NAM_DEFINE_AXPY8(16)
NAM_DEFINE_AXPY8(32)

static inline void axpy_rows_4(float* restrict acc, const float* restrict wt, const float* restrict x, int ic)
{
  v4f a = *(const v4f*)acc;
  for (int i = 0; i < ic; i++)
  {
    const float xi = x[i];
    const v4f xv = {xi, xi, xi, xi};
    a += *(const v4f*)(wt + (size_t)i * 4) * xv;
  }
  *(v4f*)acc = a;
}

static inline void axpy_rows(float* restrict acc, const float* restrict wt, const float* restrict x, int ic, int oc)
{
  switch (oc)
  {
    case 4: axpy_rows_4(acc, wt, x, ic); return;
    case 8: axpy_rows_8(acc, wt, x, ic); return;
    case 16: axpy_rows_16(acc, wt, x, ic); return;
    case 32: axpy_rows_32(acc, wt, x, ic); return;
    default:
      for (int i = 0; i < ic; i++)
      {
        const float xi = x[i];
        const float* wc = wt + (size_t)i * oc;
        for (int o = 0; o < oc; o++)
          acc[o] += wc[o] * xi;
      }
  }
}

/* in: (in_ch x n) with column stride in_stride; out: (out_ch x n) contiguous.
 * y[o] = sum_i W[o,i] x[i] accumulated in ascending i from zero, bias added last. */
#define NAM_MAX_VEC 64
static void conv1x1_process(const conv1x1_t* m, const float* in, int in_stride, float* out, int n)
{
  const int ic = m->in_ch, oc = m->out_ch;
  if (oc <= NAM_MAX_VEC)
  {
    for (int f = 0; f < n; f++)
    {
      const float* x = in + (size_t)f * in_stride;
      float* y = out + (size_t)f * oc;
      float acc[NAM_MAX_VEC];
      for (int o = 0; o < oc; o++)
        acc[o] = 0.0f;
      axpy_rows(acc, m->wt, x, ic, oc);
      if (m->has_bias)
        for (int o = 0; o < oc; o++)
          y[o] = acc[o] + m->b[o];
      else
        for (int o = 0; o < oc; o++)
          y[o] = acc[o];
    }
    return;
  }
  for (int f = 0; f < n; f++)
  {
    const float* x = in + (size_t)f * in_stride;
    float* y = out + (size_t)f * oc;
    for (int o = 0; o < oc; o++)
    {
      const float* wr = m->w + (size_t)o * ic;
      float acc = 0.0f;
      for (int i = 0; i < ic; i++)
        acc += wr[i] * x[i];
      y[o] = m->has_bias ? acc + m->b[o] : acc;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Conv1D  (NAM/conv1d.cpp:11-56 weights, :128-149 buffers, :163-183,666-683,769-774 Process;
 *          history semantics NAM/ring_buffer.cpp:7-57: zero history of (K-1)*dilation columns)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int in_ch, out_ch, kernel, dilation, groups, has_bias;
  float* w; /* [k][o][i] dense, block diagonal when grouped */
  float* wt; /* transposed copy [k][i][o] */
  float* b;
  long lookback; /* (K-1)*dilation */
  int max_buf;
  /* in_ch x (2*lookback + max_buf) column-major history with a write cursor, rewound by copying
   * the last `lookback` columns to the front when full (NAM/ring_buffer.cpp:7-27,83-109) */
  float* hist;
  long hist_cols, write_pos;
} conv1d_t;

static int conv1d_init(conv1d_t* m, int in_ch, int out_ch, int kernel, int bias, int dilation, int groups)
{
  memset(m, 0, sizeof(*m));
  if (groups <= 0 || in_ch % groups != 0 || out_ch % groups != 0)
  {
    set_err("Conv1D: channels (%d -> %d) must be divisible by groups (%d)", in_ch, out_ch, groups);
    return -1;
  }
  if (kernel < 1 || dilation < 1)
  {
    set_err("Conv1D: kernel (%d) and dilation (%d) must be >= 1", kernel, dilation);
    return -1;
  }
  m->in_ch = in_ch;
  m->out_ch = out_ch;
  m->kernel = kernel;
  m->dilation = dilation;
  m->groups = groups;
  m->has_bias = bias;
  m->lookback = (long)(kernel - 1) * dilation;
  m->w = (float*)xcalloc((size_t)kernel * out_ch * in_ch, sizeof(float));
  m->wt = (float*)xcalloc((size_t)kernel * out_ch * in_ch, sizeof(float));
  m->b = (float*)xcalloc((size_t)out_ch, sizeof(float));
  return 0;
}

/* .nam order: for g, for out-in-group i, for in-in-group j, for tap k (k innermost);
 * NAM/conv1d.cpp:40-52. */
static void conv1d_set_weights(conv1d_t* m, cursor_t* c)
{
  const int opg = m->out_ch / m->groups, ipg = m->in_ch / m->groups;
  for (int g = 0; g < m->groups; g++)
    for (int i = 0; i < opg; i++)
      for (int j = 0; j < ipg; j++)
        for (int k = 0; k < m->kernel; k++)
          m->w[((size_t)k * m->out_ch + (g * opg + i)) * m->in_ch + (g * ipg + j)] = take_w(c);
  if (m->has_bias)
    for (int i = 0; i < m->out_ch; i++)
      m->b[i] = take_w(c);
  for (int k = 0; k < m->kernel; k++)
    for (int o = 0; o < m->out_ch; o++)
      for (int i = 0; i < m->in_ch; i++)
        m->wt[((size_t)k * m->in_ch + i) * m->out_ch + o] = m->w[((size_t)k * m->out_ch + o) * m->in_ch + i];
}

static void conv1d_set_max_buffer(conv1d_t* m, int max_buf)
{
  free(m->hist);
  m->max_buf = max_buf;
  m->hist_cols = 2 * m->lookback + max_buf;
  m->hist = (float*)xcalloc((size_t)m->in_ch * (size_t)m->hist_cols, sizeof(float));
  m->write_pos = m->lookback;
}

static void conv1d_free(conv1d_t* m)
{
  free(m->w);
  free(m->wt);
  free(m->b);
  free(m->hist);
  m->w = m->wt = m->b = m->hist = NULL;
}

static void conv1d_clone(conv1d_t* d, const conv1d_t* s)
{
  *d = *s;
  const size_t nw = (size_t)s->kernel * s->out_ch * s->in_ch;
  d->w = (float*)xcalloc(nw, sizeof(float));
  memcpy(d->w, s->w, nw * sizeof(float));
  d->wt = (float*)xcalloc(nw, sizeof(float));
  memcpy(d->wt, s->wt, nw * sizeof(float));
  d->b = (float*)xcalloc((size_t)s->out_ch, sizeof(float));
  memcpy(d->b, s->b, sizeof(float) * (size_t)s->out_ch);
  if (s->hist)
  {
    const size_t nh = (size_t)s->in_ch * (size_t)s->hist_cols;
    d->hist = (float*)xcalloc(nh, sizeof(float));
    memcpy(d->hist, s->hist, nh * sizeof(float));
  }
}

/* in: (in_ch x n) contiguous; out: (out_ch x n) contiguous.
 * y[o] = b[o] + sum_k sum_i W_k[o,i] x[t-(K-1-k)d][i], taps oldest first (k = 0 is the oldest
 * sample, conv1d.cpp:678-680), input channels ascending, one running accumulator from zero. */
static void conv1d_process(conv1d_t* m, const float* in, float* out, int n)
{
  const int ic = m->in_ch, oc = m->out_ch;
  long lb = m->lookback;
  /* Write, rewinding first if the block would not fit (ring_buffer.cpp:29-42,83-109) */
  if (m->write_pos + n > m->hist_cols)
  {
    if (lb > 0)
      memmove(m->hist, m->hist + (size_t)(m->write_pos - lb) * ic, sizeof(float) * (size_t)ic * lb);
    m->write_pos = lb;
  }
  memcpy(m->hist + (size_t)m->write_pos * ic, in, sizeof(float) * (size_t)ic * n);
  lb = m->write_pos; /* column of frame 0 of this block */
  for (int f = 0; f < n; f++)
  {
    float* y = out + (size_t)f * oc;
    if (oc <= NAM_MAX_VEC)
    {
      float acc[NAM_MAX_VEC];
      for (int o = 0; o < oc; o++)
        acc[o] = 0.0f;
      for (int k = 0; k < m->kernel; k++)
      {
        const long col = lb + f - (long)(m->kernel - 1 - k) * m->dilation;
        const float* x = m->hist + (size_t)col * ic;
        axpy_rows(acc, m->wt + (size_t)k * ic * oc, x, ic, oc);
      }
      if (m->has_bias)
        for (int o = 0; o < oc; o++)
          y[o] = acc[o] + m->b[o];
      else
        for (int o = 0; o < oc; o++)
          y[o] = acc[o];
      continue;
    }
    for (int o = 0; o < oc; o++)
    {
      float acc = 0.0f;
      for (int k = 0; k < m->kernel; k++)
      {
        const long col = lb + f - (long)(m->kernel - 1 - k) * m->dilation;
        const float* x = m->hist + (size_t)col * ic;
        const float* wr = m->w + ((size_t)k * oc + o) * ic;
        for (int i = 0; i < ic; i++)
          acc += wr[i] * x[i];
      }
      y[o] = m->has_bias ? acc + m->b[o] : acc;
    }
  }
  m->write_pos += n; /* Advance (ring_buffer.cpp:59-62) */
}

/* ------------------------------------------------------------------------------------------
 * FiLM  (NAM/film.h:21-204): scale/shift = Conv1x1(cond) with bias; out = in*scale (+shift)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int active, do_shift, dim;
  conv1x1_t css;
  float* ss; /* ((shift?2:1)*dim x max_buf) */
  float* out; /* (dim x max_buf) */
} film_t;

static int film_init(film_t* f, int active, int cond_dim, int dim, int shift, int groups)
{
  memset(f, 0, sizeof(*f));
  f->active = active;
  if (!active)
    return 0;
  f->do_shift = shift;
  f->dim = dim;
  return conv1x1_init(&f->css, cond_dim, (shift ? 2 : 1) * dim, 1, groups);
}
static void film_set_max_buffer(film_t* f, int max_buf)
{
  if (!f->active)
    return;
  free(f->ss);
  free(f->out);
  f->ss = (float*)xcalloc((size_t)f->css.out_ch * max_buf, sizeof(float));
  f->out = (float*)xcalloc((size_t)f->dim * max_buf, sizeof(float));
}
static void film_free(film_t* f)
{
  if (!f->active)
    return;
  conv1x1_free(&f->css);
  free(f->ss);
  free(f->out);
}
static void film_clone(film_t* d, const film_t* s, int max_buf)
{
  *d = *s;
  if (!s->active)
    return;
  conv1x1_clone(&d->css, &s->css);
  d->ss = d->out = NULL;
  if (s->ss)
    film_set_max_buffer(d, max_buf);
}
/* Process: out (dim x n, contiguous in f->out) from in (dim rows, column stride in_stride) */
static void film_process(film_t* f, const float* in, int in_stride, const float* cond, int cond_dim, int n)
{
  conv1x1_process(&f->css, cond, cond_dim, f->ss, n);
  const int d = f->dim, rows = f->css.out_ch;
  for (int c = 0; c < n; c++)
  {
    const float* x = in + (size_t)c * in_stride;
    const float* sc = f->ss + (size_t)c * rows;
    float* y = f->out + (size_t)c * d;
    if (f->do_shift)
      for (int i = 0; i < d; i++)
        y[i] = x[i] * sc[i] + sc[d + i];
    else
      for (int i = 0; i < d; i++)
        y[i] = x[i] * sc[i];
  }
}
/* Process_: in place (film.h:192-197) */
static void film_process_inplace(film_t* f, float* io, int stride, const float* cond, int cond_dim, int n)
{
  film_process(f, io, stride, cond, cond_dim, n);
  for (int c = 0; c < n; c++)
    memcpy(io + (size_t)c * stride, f->out + (size_t)c * f->dim, sizeof(float) * (size_t)f->dim);
}

/* ------------------------------------------------------------------------------------------
 * WaveNet Layer  (NAM/wavenet/detail.h:44-158 construction, NAM/wavenet/model.cpp:152-181
 * weight order, :183-393 Process)
 * ---------------------------------------------------------------------------------------- */
enum { GATING_NONE = 0, GATING_GATED = 1, GATING_BLENDED = 2 };
enum { F_CONV_PRE = 0, F_CONV_POST, F_MIXIN_PRE, F_MIXIN_POST, F_ACT_PRE, F_ACT_POST, F_L1X1_POST, F_H1X1_POST, F_COUNT };

typedef struct
{
  int channels, bottleneck, condition_size, gating;
  conv1d_t conv;
  conv1x1_t mixin;
  int has_l1x1, has_h1x1;
  conv1x1_t l1x1, h1x1;
  act_t act, sec_act;
  film_t film[F_COUNT];
  /* scratch, sized in set_max_buffer */
  float *conv_out, *mixin_out, *z, *l1x1_out, *h1x1_out, *out_next, *out_head, *zc;
  int zrows, head_rows;
} layer_t;

typedef struct
{
  int input_size, condition_size, channels, bottleneck, head_size, head_kernel, head_dilation, head_bias;
  int groups_input, groups_mixin, l1x1_active, l1x1_groups, h1x1_active, h1x1_out, h1x1_groups;
  int film_active[F_COUNT], film_shift[F_COUNT], film_groups[F_COUNT];
} array_params_t;

static int layer_init(layer_t* L, const array_params_t* p, int kernel, int dilation, int gating, cursor_t* c)
{
  memset(L, 0, sizeof(*L));
  L->channels = p->channels;
  L->bottleneck = p->bottleneck;
  L->condition_size = p->condition_size;
  L->gating = gating;
  const int zrows = gating != GATING_NONE ? 2 * p->bottleneck : p->bottleneck;
  L->zrows = zrows;
  if (conv1d_init(&L->conv, p->channels, zrows, kernel, 1, dilation, p->groups_input))
    return -1;
  if (conv1x1_init(&L->mixin, p->condition_size, zrows, 0, p->groups_mixin))
    return -1;
  act_parse(&L->act, c);
  act_parse(&L->sec_act, c);
  L->has_l1x1 = p->l1x1_active;
  if (p->l1x1_active)
  {
    if (conv1x1_init(&L->l1x1, p->bottleneck, p->channels, 1, p->l1x1_groups))
      return -1;
  }
  else
  {
    if (p->bottleneck != p->channels)
    {
      set_err("When layer1x1.active is false, bottleneck (%d) must equal channels (%d)", p->bottleneck, p->channels);
      return -1;
    }
    if (p->film_active[F_L1X1_POST])
    {
      set_err("layer1x1_post_film cannot be active when layer1x1 is not active");
      return -1;
    }
  }
  L->has_h1x1 = p->h1x1_active;
  if (p->h1x1_active)
  {
    if (conv1x1_init(&L->h1x1, p->bottleneck, p->h1x1_out, 1, p->h1x1_groups))
      return -1;
  }
  else if (p->film_active[F_H1X1_POST])
  {
    set_err("Do not use post-head 1x1 FiLM if there is no head 1x1");
    return -1;
  }
  L->head_rows = p->h1x1_active ? p->h1x1_out : p->bottleneck;
  /* FiLM dims: detail.h:104-157 */
  const int dims[F_COUNT] = {p->channels, zrows, p->condition_size, zrows, zrows, p->bottleneck, p->channels,
                             p->h1x1_out};
  for (int i = 0; i < F_COUNT; i++)
  {
    int active = p->film_active[i];
    if (i == F_L1X1_POST && !p->l1x1_active)
      active = 0;
    if (i == F_H1X1_POST && !p->h1x1_active)
      active = 0;
    if (film_init(&L->film[i], active, p->condition_size, dims[i], p->film_shift[i], p->film_groups[i]))
      return -1;
  }
  return 0;
}

/* model.cpp:152-181 */
static void layer_set_weights(layer_t* L, cursor_t* c)
{
  conv1d_set_weights(&L->conv, c);
  conv1x1_set_weights(&L->mixin, c);
  if (L->has_l1x1)
    conv1x1_set_weights(&L->l1x1, c);
  if (L->has_h1x1)
    conv1x1_set_weights(&L->h1x1, c);
  for (int i = 0; i < F_COUNT; i++)
    if (L->film[i].active)
      conv1x1_set_weights(&L->film[i].css, c);
}

static void layer_free_scratch(layer_t* L)
{
  free(L->conv_out);
  free(L->mixin_out);
  free(L->z);
  free(L->l1x1_out);
  free(L->h1x1_out);
  free(L->out_next);
  free(L->out_head);
  free(L->zc);
  L->conv_out = L->mixin_out = L->z = L->l1x1_out = L->h1x1_out = L->out_next = L->out_head = L->zc = NULL;
}

static void layer_set_max_buffer(layer_t* L, int mb)
{
  layer_free_scratch(L);
  conv1d_set_max_buffer(&L->conv, mb);
  L->conv_out = (float*)xcalloc((size_t)L->zrows * mb, sizeof(float));
  L->mixin_out = (float*)xcalloc((size_t)L->zrows * mb, sizeof(float));
  L->z = (float*)xcalloc((size_t)L->zrows * mb, sizeof(float));
  L->zc = (float*)xcalloc((size_t)L->bottleneck * mb, sizeof(float));
  L->l1x1_out = (float*)xcalloc((size_t)L->channels * mb, sizeof(float));
  L->h1x1_out = (float*)xcalloc((size_t)(L->has_h1x1 ? L->h1x1.out_ch : 1) * mb, sizeof(float));
  L->out_next = (float*)xcalloc((size_t)L->channels * mb, sizeof(float));
  L->out_head = (float*)xcalloc((size_t)L->head_rows * mb, sizeof(float));
  for (int i = 0; i < F_COUNT; i++)
    film_set_max_buffer(&L->film[i], mb);
}

static void layer_free(layer_t* L)
{
  conv1d_free(&L->conv);
  conv1x1_free(&L->mixin);
  if (L->has_l1x1)
    conv1x1_free(&L->l1x1);
  if (L->has_h1x1)
    conv1x1_free(&L->h1x1);
  act_free(&L->act);
  act_free(&L->sec_act);
  for (int i = 0; i < F_COUNT; i++)
    film_free(&L->film[i]);
  layer_free_scratch(L);
}

static void layer_clone(layer_t* d, const layer_t* s)
{
  *d = *s;
  conv1d_clone(&d->conv, &s->conv);
  conv1x1_clone(&d->mixin, &s->mixin);
  if (s->has_l1x1)
    conv1x1_clone(&d->l1x1, &s->l1x1);
  if (s->has_h1x1)
    conv1x1_clone(&d->h1x1, &s->h1x1);
  act_clone(&d->act, &s->act);
  act_clone(&d->sec_act, &s->sec_act);
  for (int i = 0; i < F_COUNT; i++)
    film_clone(&d->film[i], &s->film[i], s->conv.max_buf);
  d->conv_out = d->mixin_out = d->z = d->l1x1_out = d->h1x1_out = d->out_next = d->out_head = d->zc = NULL;
  if (s->z)
  {
    const int mb = s->conv.max_buf;
    /* conv1d_clone already copied the history; re-create scratch without touching it */
    d->conv_out = (float*)xcalloc((size_t)d->zrows * mb, sizeof(float));
    d->mixin_out = (float*)xcalloc((size_t)d->zrows * mb, sizeof(float));
    d->z = (float*)xcalloc((size_t)d->zrows * mb, sizeof(float));
    d->zc = (float*)xcalloc((size_t)d->bottleneck * mb, sizeof(float));
    d->l1x1_out = (float*)xcalloc((size_t)d->channels * mb, sizeof(float));
    d->h1x1_out = (float*)xcalloc((size_t)(d->has_h1x1 ? d->h1x1.out_ch : 1) * mb, sizeof(float));
    d->out_next = (float*)xcalloc((size_t)d->channels * mb, sizeof(float));
    d->out_head = (float*)xcalloc((size_t)d->head_rows * mb, sizeof(float));
  }
}

/* Gating / blending on a (2B x n) block, result in the top B rows of each column
 * (gating_activations.h:100-113, :209-227; model.cpp:252-287). */
static void gate_apply(const act_t* act, const act_t* sec, int mode, int B, float* z, int zrows, int n, int fast_tanh)
{
  float a[256], g[256];
  for (int f = 0; f < n; f++)
  {
    float* col = z + (size_t)f * zrows;
    int done = 0;
    while (done < B)
    {
      /* process in chunks of <= 256 channels; PReLU indexing needs absolute channel */
      const int m = (B - done) < 256 ? (B - done) : 256;
      if (done == 0 && m == B)
      {
        memcpy(a, col, sizeof(float) * (size_t)B);
        memcpy(g, col + B, sizeof(float) * (size_t)B);
        act_apply(act, a, B, fast_tanh);
        act_apply(sec, g, B, fast_tanh);
        if (mode == GATING_GATED)
          for (int c = 0; c < B; c++)
            col[c] = a[c] * g[c];
        else
          for (int c = 0; c < B; c++)
            col[c] = g[c] * a[c] + (1.0f - g[c]) * col[c];
      }
      else
      {
        /* bottleneck > 256 never occurs in NAM models; fail loudly rather than mis-index PReLU */
        fprintf(stderr, "nam_oracle: gating with bottleneck > 256 unsupported\n");
        abort();
      }
      done += m;
    }
  }
}

/* model.cpp:183-393.  input (channels x n), cond (condition_size x n), both contiguous. */
static void layer_process(layer_t* L, const float* input, const float* cond, int n, int fast_tanh)
{
  const int C = L->channels, B = L->bottleneck, Z = L->zrows, cs = L->condition_size;
  /* Step 1: input convolution with optional pre/post FiLM (:189-204) */
  if (L->film[F_CONV_PRE].active)
  {
    film_process(&L->film[F_CONV_PRE], input, C, cond, cs, n);
    conv1d_process(&L->conv, L->film[F_CONV_PRE].out, L->conv_out, n);
  }
  else
    conv1d_process(&L->conv, input, L->conv_out, n);
  if (L->film[F_CONV_POST].active)
    film_process_inplace(&L->film[F_CONV_POST], L->conv_out, Z, cond, cs, n);
  /* input mixin (:206-219) */
  if (L->film[F_MIXIN_PRE].active)
  {
    film_process(&L->film[F_MIXIN_PRE], cond, cs, cond, cs, n);
    conv1x1_process(&L->mixin, L->film[F_MIXIN_PRE].out, cs, L->mixin_out, n);
  }
  else
    conv1x1_process(&L->mixin, cond, cs, L->mixin_out, n);
  if (L->film[F_MIXIN_POST].active)
    film_process_inplace(&L->film[F_MIXIN_POST], L->mixin_out, Z, cond, cs, n);
  /* z = conv + mixin (:220-221) */
  for (long p = 0; p < (long)Z * n; p++)
    L->z[p] = L->conv_out[p] + L->mixin_out[p];
  if (L->film[F_ACT_PRE].active)
    film_process_inplace(&L->film[F_ACT_PRE], L->z, Z, cond, cs, n);

  /* Steps 2 & 3: activation and layer1x1 (:234-288).  `za` points at the activated
   * (bottleneck x n) block with column stride `zs`. */
  const float* za = L->z;
  int zs = Z;
  if (L->gating == GATING_NONE)
  {
    act_apply(&L->act, L->z, (long)Z * n, fast_tanh);
    if (L->film[F_ACT_POST].active)
      film_process_inplace(&L->film[F_ACT_POST], L->z, Z, cond, cs, n);
    if (L->has_l1x1)
      conv1x1_process(&L->l1x1, L->z, Z, L->l1x1_out, n);
  }
  else
  {
    gate_apply(&L->act, &L->sec_act, L->gating, B, L->z, Z, n, fast_tanh);
    if (L->film[F_ACT_POST].active)
      film_process_inplace(&L->film[F_ACT_POST], L->z, Z, cond, cs, n); /* operates on top B rows, stride Z */
    if (L->has_l1x1)
    {
      conv1x1_process(&L->l1x1, L->z, Z, L->l1x1_out, n);
      /* Reference quirk: layer1x1_post_film is applied only in BLENDED mode (:279-287) */
      if (L->gating == GATING_BLENDED && L->film[F_L1X1_POST].active)
        film_process_inplace(&L->film[F_L1X1_POST], L->l1x1_out, C, cond, cs, n);
    }
  }

  /* head output (:290-352) */
  if (L->has_h1x1)
  {
    conv1x1_process(&L->h1x1, za, zs, L->h1x1_out, n);
    if (L->film[F_H1X1_POST].active)
      film_process_inplace(&L->film[F_H1X1_POST], L->h1x1_out, L->h1x1.out_ch, cond, cs, n);
    memcpy(L->out_head, L->h1x1_out, sizeof(float) * (size_t)L->h1x1.out_ch * n);
  }
  else
  {
    for (int f = 0; f < n; f++)
      memcpy(L->out_head + (size_t)f * B, za + (size_t)f * zs, sizeof(float) * (size_t)B);
  }

  /* residual (:354-392) */
  if (L->has_l1x1)
    for (long p = 0; p < (long)C * n; p++)
      L->out_next[p] = input[p] + L->l1x1_out[p];
  else
    memcpy(L->out_next, input, sizeof(float) * (size_t)C * n);
}

/* ------------------------------------------------------------------------------------------
 * LayerArray  (model.cpp:397-575)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  array_params_t p;
  int n_layers, head_out_size;
  conv1x1_t rechannel;
  layer_t* layers;
  conv1d_t head_rechannel;
  float *rech_out, *head_inputs, *head_out;
} larray_t;

static void larray_free(larray_t* A)
{
  conv1x1_free(&A->rechannel);
  for (int i = 0; i < A->n_layers; i++)
    layer_free(&A->layers[i]);
  free(A->layers);
  conv1d_free(&A->head_rechannel);
  free(A->rech_out);
  free(A->head_inputs);
  free(A->head_out);
}

static void larray_set_max_buffer(larray_t* A, int mb)
{
  free(A->rech_out);
  free(A->head_inputs);
  free(A->head_out);
  A->rech_out = (float*)xcalloc((size_t)A->p.channels * mb, sizeof(float));
  A->head_inputs = (float*)xcalloc((size_t)A->head_out_size * mb, sizeof(float));
  A->head_out = (float*)xcalloc((size_t)A->p.head_size * mb, sizeof(float));
  conv1d_set_max_buffer(&A->head_rechannel, mb);
  for (int i = 0; i < A->n_layers; i++)
    layer_set_max_buffer(&A->layers[i], mb);
}

static long larray_receptive_field(const larray_t* A)
{
  long r = 0;
  for (int i = 0; i < A->n_layers; i++)
    r += (long)A->layers[i].conv.dilation * (A->layers[i].conv.kernel - 1);
  r += (long)A->head_rechannel.dilation * (A->head_rechannel.kernel - 1);
  return r;
}

/* ProcessInner (model.cpp:488-549). head_inputs must already hold the initial accumulator.
 * Returns pointer to the last layer's residual output (channels x n). */
static const float* larray_process(larray_t* A, const float* layer_inputs, const float* cond, int n, int fast_tanh)
{
  conv1x1_process(&A->rechannel, layer_inputs, A->p.input_size, A->rech_out, n);
  const float* x = A->rech_out;
  for (int i = 0; i < A->n_layers; i++)
  {
    layer_process(&A->layers[i], x, cond, n, fast_tanh);
    const float* h = A->layers[i].out_head;
    for (long p = 0; p < (long)A->head_out_size * n; p++)
      A->head_inputs[p] += h[p];
    x = A->layers[i].out_next;
  }
  conv1d_process(&A->head_rechannel, A->head_inputs, A->head_out, n);
  return x;
}

/* ------------------------------------------------------------------------------------------
 * Post-stack Head (model.cpp:19-103): repeated (activation -> Conv1D, dilation 1, bias)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int n, in_ch, out_ch;
  conv1d_t* convs;
  act_t act;
  float** outs;
  float* work;
} pshead_t;

/* ------------------------------------------------------------------------------------------
 * LSTM (NAM/lstm.cpp:9-29 weights, :31-68 cell, :103-168 process)
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int input_size, hidden;
  float* w; /* (4H x (I+H)) row-major */
  float* b; /* 4H */
  float* xh; /* I+H : [x ; h] */
  float* c; /* H */
  float* ifgo; /* 4H scratch */
} lstm_cell_t;

/* ------------------------------------------------------------------------------------------
 * The model object
 * ---------------------------------------------------------------------------------------- */
struct nam_oracle
{
  int arch, in_ch, out_ch, fast_tanh;
  double expected_sr;
  int max_buf, prewarm_samples, weights_consumed;
  /* wavenet */
  int n_arrays, cond_dim;
  larray_t* arrays;
  float head_scale;
  int with_head;
  pshead_t ph;
  struct nam_oracle* cond_dsp;
  float *cond_in, *cond_out;
  /* lstm */
  int n_cells, lstm_input, lstm_hidden;
  lstm_cell_t* cells;
  float *head_w, *head_b;
  /* linear */
  int rf, lin_bias;
  float* lin_w; /* reversed impulse response (NAM/linear.cpp:71-74) */
  float lin_b;
  float* lin_hist; /* in_ch x (rf + max_buf) row per channel */
  /* convnet (NAM/convnet.cpp) */
  int cn_blocks, cn_channels, cn_batchnorm;
  conv1d_t* cn_convs;
  float *cn_scale, *cn_loc; /* [block][channels] */
  act_t cn_act;
  float *cn_head_w, *cn_head_b;
  float* cn_buf[2]; /* (channels | in_ch) x max_buf, frame-major */
};

static int parse_array_params(array_params_t* p, cursor_t* c)
{
  p->input_size = take_i(c);
  p->condition_size = take_i(c);
  p->channels = take_i(c);
  p->bottleneck = take_i(c);
  p->head_size = take_i(c);
  p->head_kernel = take_i(c);
  p->head_dilation = take_i(c);
  p->head_bias = take_i(c);
  p->groups_input = take_i(c);
  p->groups_mixin = take_i(c);
  p->l1x1_active = take_i(c);
  p->l1x1_groups = take_i(c);
  p->h1x1_active = take_i(c);
  p->h1x1_out = take_i(c);
  p->h1x1_groups = take_i(c);
  for (int i = 0; i < F_COUNT; i++)
  {
    p->film_active[i] = take_i(c);
    p->film_shift[i] = take_i(c);
    p->film_groups[i] = take_i(c);
  }
  if (p->head_kernel < 1)
  {
    set_err("head.kernel_size must be >= 1");
    return -1;
  }
  return 0;
}

static int build_wavenet(nam_oracle* o, cursor_t* c)
{
  o->in_ch = take_i(c);
  o->n_arrays = take_i(c);
  o->with_head = take_i(c);
  (void)take_f(c); /* JSON head_scale: overridden by the last weight (model.cpp:670) */
  if (o->n_arrays <= 0)
  {
    set_err("WaveNet requires at least one layer array");
    return -1;
  }
  o->arrays = (larray_t*)xcalloc((size_t)o->n_arrays, sizeof(larray_t));
  for (int a = 0; a < o->n_arrays; a++)
  {
    larray_t* A = &o->arrays[a];
    if (parse_array_params(&A->p, c))
      return -1;
    A->head_out_size = A->p.h1x1_active ? A->p.h1x1_out : A->p.bottleneck;
    if (conv1x1_init(&A->rechannel, A->p.input_size, A->p.channels, 0, 1))
      return -1;
    /* model.cpp:399-400: Conv1D(head_out_size -> head_size, head_kernel, head_bias, head_dilation) */
    if (conv1d_init(&A->head_rechannel, A->head_out_size, A->p.head_size, A->p.head_kernel, A->p.head_bias,
                    A->p.head_dilation, 1))
      return -1;
    A->n_layers = take_i(c);
    if (A->n_layers <= 0)
    {
      set_err("layer array %d has no layers", a);
      return -1;
    }
    A->layers = (layer_t*)xcalloc((size_t)A->n_layers, sizeof(layer_t));
    for (int l = 0; l < A->n_layers; l++)
    {
      const int kernel = take_i(c), dilation = take_i(c), gating = take_i(c);
      if (layer_init(&A->layers[l], &A->p, kernel, dilation, gating, c))
        return -1;
    }
    if (a > 0)
    {
      /* model.cpp:639-647 */
      if (A->p.channels != o->arrays[a - 1].p.head_size)
      {
        set_err("channels of layer %d (%d) doesn't match head_size of preceding layer (%d)", a, A->p.channels,
                o->arrays[a - 1].p.head_size);
        return -1;
      }
      /* the head accumulator is initialised by a straight copy (model.cpp:473-486) */
      if (A->head_out_size != o->arrays[a - 1].p.head_size)
      {
        set_err("layer array %d: head accumulator rows (%d) != previous head_size (%d)", a, A->head_out_size,
                o->arrays[a - 1].p.head_size);
        return -1;
      }
    }
  }
  o->cond_dim = o->in_ch; /* _get_condition_dim(), model.h */
  if (o->cond_dsp)
  {
    if (o->cond_dsp->in_ch != o->cond_dim)
    {
      set_err("input channels of WaveNet (%d) don't match input channels of condition DSP (%d)", o->in_ch,
              o->cond_dsp->in_ch);
      return -1;
    }
    for (int a = 0; a < o->n_arrays; a++)
      if (o->arrays[a].p.condition_size != o->cond_dsp->out_ch)
      {
        set_err("condition_size of layer %d (%d) doesn't match output channels of condition DSP (%d)", a,
                o->arrays[a].p.condition_size, o->cond_dsp->out_ch);
        return -1;
      }
    if (o->cond_dsp->expected_sr != o->expected_sr)
    {
      set_err("Condition DSP expected sample rate (%g) doesn't match WaveNet expected sample rate (%g)",
              o->cond_dsp->expected_sr, o->expected_sr);
      return -1;
    }
  }
  o->out_ch = o->arrays[o->n_arrays - 1].p.head_size;
  if (o->with_head)
  {
    pshead_t* H = &o->ph;
    const int channels = take_i(c), out_channels = take_i(c), nk = take_i(c);
    if (nk <= 0)
    {
      set_err("WaveNet Head: kernel_sizes must be non-empty");
      return -1;
    }
    H->n = nk;
    H->in_ch = o->out_ch;
    H->out_ch = out_channels;
    H->convs = (conv1d_t*)xcalloc((size_t)nk, sizeof(conv1d_t));
    H->outs = (float**)xcalloc((size_t)nk, sizeof(float*));
    int cin = H->in_ch;
    for (int i = 0; i < nk; i++)
    {
      const int k = take_i(c);
      const int cout = (i + 1 == nk) ? out_channels : channels;
      if (k < 1)
      {
        set_err("WaveNet Head: kernel_sizes entries must be >= 1");
        return -1;
      }
      if (conv1d_init(&H->convs[i], cin, cout, k, 1, 1, 1))
        return -1;
      cin = cout;
    }
    act_parse(&H->act, c);
    o->out_ch = out_channels;
  }
  if (c->failed)
  {
    set_err("config stream truncated");
    return -1;
  }
  /* set_weights_ (model.cpp:661-683) */
  for (int a = 0; a < o->n_arrays; a++)
  {
    larray_t* A = &o->arrays[a];
    conv1x1_set_weights(&A->rechannel, c);
    for (int l = 0; l < A->n_layers; l++)
      layer_set_weights(&A->layers[l], c);
    conv1d_set_weights(&A->head_rechannel, c);
  }
  if (o->with_head)
    for (int i = 0; i < o->ph.n; i++)
      conv1d_set_weights(&o->ph.convs[i], c);
  o->head_scale = take_w(c);
  if (c->failed == 2)
  {
    set_err("Weight mismatch: provided %d weights, but the model expects more.", c->n_w);
    return -1;
  }
  if (c->i_w != c->n_w)
  {
    set_err("Weight mismatch: assigned %d weights, but %d were provided.", c->i_w, c->n_w);
    return -1;
  }
  /* prewarm (model.cpp:653-658) */
  o->prewarm_samples = o->cond_dsp ? o->cond_dsp->prewarm_samples : 1;
  for (int a = 0; a < o->n_arrays; a++)
    o->prewarm_samples += (int)larray_receptive_field(&o->arrays[a]);
  if (o->with_head)
  {
    long rf = 1;
    for (int i = 0; i < o->ph.n; i++)
      rf += o->ph.convs[i].kernel - 1;
    o->prewarm_samples += (int)rf - 1;
  }
  return 0;
}

static int build_lstm(nam_oracle* o, cursor_t* c)
{
  o->in_ch = take_i(c);
  o->out_ch = take_i(c);
  o->n_cells = take_i(c);
  o->lstm_input = take_i(c);
  o->lstm_hidden = take_i(c);
  if (c->failed || o->n_cells < 0 || o->lstm_hidden <= 0 || o->lstm_input <= 0)
  {
    set_err("bad LSTM config");
    return -1;
  }
  const int H = o->lstm_hidden;
  o->cells = (lstm_cell_t*)xcalloc((size_t)o->n_cells, sizeof(lstm_cell_t));
  for (int l = 0; l < o->n_cells; l++)
  {
    lstm_cell_t* L = &o->cells[l];
    const int I = l == 0 ? o->lstm_input : H;
    L->input_size = I;
    L->hidden = H;
    L->w = (float*)xcalloc((size_t)4 * H * (I + H), sizeof(float));
    L->b = (float*)xcalloc((size_t)4 * H, sizeof(float));
    L->xh = (float*)xcalloc((size_t)(I + H), sizeof(float));
    L->c = (float*)xcalloc((size_t)H, sizeof(float));
    L->ifgo = (float*)xcalloc((size_t)4 * H, sizeof(float));
    for (int i = 0; i < 4 * H * (I + H); i++)
      L->w[i] = take_w(c);
    for (int i = 0; i < 4 * H; i++)
      L->b[i] = take_w(c);
    for (int i = 0; i < H; i++)
      L->xh[I + i] = take_w(c); /* initial hidden state is a trained parameter (lstm.cpp:24-26) */
    for (int i = 0; i < H; i++)
      L->c[i] = take_w(c);
  }
  o->head_w = (float*)xcalloc((size_t)o->out_ch * H, sizeof(float));
  o->head_b = (float*)xcalloc((size_t)o->out_ch, sizeof(float));
  for (int i = 0; i < o->out_ch * H; i++)
    o->head_w[i] = take_w(c);
  for (int i = 0; i < o->out_ch; i++)
    o->head_b[i] = take_w(c);
  if (c->failed == 2 || c->i_w != c->n_w)
  {
    set_err("LSTM weight count mismatch: consumed %d of %d", c->i_w, c->n_w);
    return -1;
  }
  /* lstm.cpp:127-134 */
  int pw = (int)(0.5 * o->expected_sr);
  o->prewarm_samples = pw <= 0 ? 1 : pw;
  return 0;
}

static int build_linear(nam_oracle* o, cursor_t* c)
{
  o->in_ch = take_i(c);
  o->out_ch = take_i(c);
  o->rf = take_i(c);
  o->lin_bias = take_i(c);
  if (c->failed || o->rf <= 0)
  {
    set_err("bad Linear config");
    return -1;
  }
  /* linear.cpp:61-81 */
  if (c->n_w != o->rf + (o->lin_bias ? 1 : 0))
  {
    set_err("Params vector does not match expected size based on architecture parameters");
    return -1;
  }
  o->lin_w = (float*)xcalloc((size_t)o->rf, sizeof(float));
  for (int i = 0; i < o->rf; i++)
    o->lin_w[i] = c->w[o->rf - 1 - i];
  o->lin_b = o->lin_bias ? c->w[o->rf] : 0.0f;
  c->i_w = c->n_w;
  o->prewarm_samples = 0; /* DSP::GetPrewarmSamples default (dsp.h:157) */
  return 0;
}

/* ConvNet (NAM/convnet.cpp:172-201 constructor, :48-60 block weights, :14-37 BatchNorm, :132-153 head).
 * cfg: in_channels, out_channels, channels, n_blocks, dilations..., batchnorm, groups, ACT */
static int build_convnet(nam_oracle* o, cursor_t* c)
{
  o->in_ch = take_i(c);
  o->out_ch = take_i(c);
  o->cn_channels = take_i(c);
  o->cn_blocks = take_i(c);
  if (c->failed || o->cn_blocks <= 0 || o->cn_blocks > 4096 || o->cn_channels <= 0)
  {
    set_err("bad ConvNet config");
    return -1;
  }
  int* dil = (int*)xcalloc((size_t)o->cn_blocks, sizeof(int));
  for (int i = 0; i < o->cn_blocks; i++)
    dil[i] = take_i(c);
  o->cn_batchnorm = take_i(c);
  const int groups = take_i(c);
  act_parse(&o->cn_act, c);
  const int C = o->cn_channels;
  o->cn_convs = (conv1d_t*)xcalloc((size_t)o->cn_blocks, sizeof(conv1d_t));
  o->cn_scale = (float*)xcalloc((size_t)o->cn_blocks * C, sizeof(float));
  o->cn_loc = (float*)xcalloc((size_t)o->cn_blocks * C, sizeof(float));
  o->prewarm_samples = 1; /* convnet.cpp:198-200 */
  for (int i = 0; i < o->cn_blocks; i++)
  {
    /* "HACK 2 kernel" (convnet.cpp:55-56): kernel size 2, bias only without batchnorm */
    if (conv1d_init(&o->cn_convs[i], i == 0 ? o->in_ch : C, C, 2, !o->cn_batchnorm, dil[i], groups) != 0)
    {
      free(dil);
      return -1;
    }
    conv1d_set_weights(&o->cn_convs[i], c);
    if (o->cn_batchnorm)
    {
      /* running_mean, running_var, weight, bias, eps -> scale = w / sqrt(eps + var), loc = b - scale * mean */
      float* tmp = (float*)xcalloc((size_t)4 * C, sizeof(float));
      for (int j = 0; j < 4 * C; j++)
        tmp[j] = take_w(c);
      const float eps = take_w(c);
      for (int j = 0; j < C; j++)
      {
        const float sc = tmp[2 * C + j] / sqrtf(eps + tmp[C + j]);
        o->cn_scale[(size_t)i * C + j] = sc;
        o->cn_loc[(size_t)i * C + j] = tmp[3 * C + j] - sc * tmp[j];
      }
      free(tmp);
    }
    o->prewarm_samples += dil[i];
  }
  free(dil);
  o->cn_head_w = (float*)xcalloc((size_t)o->out_ch * C, sizeof(float));
  o->cn_head_b = (float*)xcalloc((size_t)o->out_ch, sizeof(float));
  for (int j = 0; j < o->out_ch * C; j++)
    o->cn_head_w[j] = take_w(c);
  for (int j = 0; j < o->out_ch; j++)
    o->cn_head_b[j] = take_w(c);
  if (c->failed || c->i_w != c->n_w)
  {
    set_err("Didn't touch all the weights when initializing ConvNet"); /* convnet.cpp:194-195 */
    return -1;
  }
  return 0;
}

nam_oracle* nam_oracle_create(const int32_t* cfg, int n_cfg, const float* fparams, int n_fparams,
                              const float* weights, int n_weights, double expected_sample_rate, int fast_tanh,
                              nam_oracle* condition_dsp)
{
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.cfg = cfg;
  c.n_cfg = n_cfg;
  c.fp = fparams;
  c.n_fp = n_fparams;
  c.w = weights;
  c.n_w = n_weights;
  nam_oracle* o = (nam_oracle*)xcalloc(1, sizeof(nam_oracle));
  o->expected_sr = expected_sample_rate;
  o->fast_tanh = fast_tanh;
  o->cond_dsp = condition_dsp;
  o->arch = take_i(&c);
  int rc = -1;
  if (o->arch == NAM_ORACLE_ARCH_WAVENET)
    rc = build_wavenet(o, &c);
  else if (o->arch == NAM_ORACLE_ARCH_LSTM)
    rc = build_lstm(o, &c);
  else if (o->arch == NAM_ORACLE_ARCH_LINEAR)
    rc = build_linear(o, &c);
  else if (o->arch == NAM_ORACLE_ARCH_CONVNET)
    rc = build_convnet(o, &c);
  else
    set_err("No config parser registered for architecture code %d", o->arch);
  if (rc == 0 && (o->in_ch <= 0 || o->out_ch <= 0))
  {
    set_err("Channel counts must be positive"); /* dsp.cpp:61-64 */
    rc = -1;
  }
  if (rc != 0)
  {
    o->cond_dsp = NULL; /* caller keeps ownership on failure */
    nam_oracle_destroy(o);
    return NULL;
  }
  o->weights_consumed = c.i_w;
  return o;
}

void nam_oracle_destroy(nam_oracle* o)
{
  if (!o)
    return;
  if (o->arrays)
  {
    for (int a = 0; a < o->n_arrays; a++)
      larray_free(&o->arrays[a]);
    free(o->arrays);
  }
  if (o->ph.convs)
  {
    for (int i = 0; i < o->ph.n; i++)
    {
      conv1d_free(&o->ph.convs[i]);
      if (o->ph.outs)
        free(o->ph.outs[i]);
    }
    free(o->ph.convs);
    free(o->ph.outs);
    free(o->ph.work);
    act_free(&o->ph.act);
  }
  if (o->cond_dsp)
    nam_oracle_destroy(o->cond_dsp);
  free(o->cond_in);
  free(o->cond_out);
  if (o->cells)
  {
    for (int l = 0; l < o->n_cells; l++)
    {
      free(o->cells[l].w);
      free(o->cells[l].b);
      free(o->cells[l].xh);
      free(o->cells[l].c);
      free(o->cells[l].ifgo);
    }
    free(o->cells);
  }
  free(o->head_w);
  free(o->head_b);
  free(o->lin_w);
  free(o->lin_hist);
  if (o->cn_convs)
  {
    for (int i = 0; i < o->cn_blocks; i++)
      conv1d_free(&o->cn_convs[i]);
    free(o->cn_convs);
    act_free(&o->cn_act);
  }
  free(o->cn_scale);
  free(o->cn_loc);
  free(o->cn_head_w);
  free(o->cn_head_b);
  free(o->cn_buf[0]);
  free(o->cn_buf[1]);
  free(o);
}

int nam_oracle_in_channels(const nam_oracle* o)
{
  return o->in_ch;
}
int nam_oracle_out_channels(const nam_oracle* o)
{
  return o->out_ch;
}
int nam_oracle_prewarm_samples(const nam_oracle* o)
{
  return o->prewarm_samples;
}
int nam_oracle_weights_consumed(const nam_oracle* o)
{
  return o->weights_consumed;
}

/* SetMaxBufferSize: (re)allocate and ZERO every history (model.cpp:685-728, conv1d.cpp:128-149,
 * dsp.cpp:203-234 for Buffer). */
static void set_max_buffer(nam_oracle* o, int mb)
{
  o->max_buf = mb;
  if (o->arch == NAM_ORACLE_ARCH_WAVENET)
  {
    free(o->cond_in);
    free(o->cond_out);
    const int cond_out_dim = o->cond_dsp ? o->cond_dsp->out_ch : o->cond_dim;
    o->cond_in = (float*)xcalloc((size_t)o->cond_dim * mb, sizeof(float));
    o->cond_out = (float*)xcalloc((size_t)cond_out_dim * mb, sizeof(float));
    if (o->cond_dsp)
      set_max_buffer(o->cond_dsp, mb);
    for (int a = 0; a < o->n_arrays; a++)
      larray_set_max_buffer(&o->arrays[a], mb);
    if (o->with_head)
    {
      pshead_t* H = &o->ph;
      free(H->work);
      H->work = (float*)xcalloc((size_t)H->in_ch * mb, sizeof(float));
      for (int i = 0; i < H->n; i++)
      {
        conv1d_set_max_buffer(&H->convs[i], mb);
        free(H->outs[i]);
        H->outs[i] = (float*)xcalloc((size_t)H->convs[i].out_ch * mb, sizeof(float));
      }
    }
  }
  else if (o->arch == NAM_ORACLE_ARCH_LINEAR)
  {
    /* NOTE: the reference's Linear does NOT clear its input history on Reset (Buffer keeps
     * _input_buffers, dsp.cpp:215-233 only runs at construction); a fresh model starts from
     * zeros, which is what this restates.  Re-Reset after audio is not exercised by tests. */
    free(o->lin_hist);
    o->lin_hist = (float*)xcalloc((size_t)o->in_ch * (size_t)(o->rf + mb), sizeof(float));
  }
  else if (o->arch == NAM_ORACLE_ARCH_CONVNET)
  {
    /* ConvNet::SetMaxBufferSize (convnet.cpp:283-292): every block's Conv1D restarts from a zero history */
    const int wmax = o->cn_channels > o->in_ch ? o->cn_channels : o->in_ch;
    for (int i = 0; i < o->cn_blocks; i++)
      conv1d_set_max_buffer(&o->cn_convs[i], mb);
    for (int k = 0; k < 2; k++)
    {
      free(o->cn_buf[k]);
      o->cn_buf[k] = (float*)xcalloc((size_t)wmax * mb, sizeof(float));
    }
  }
}

static void lstm_cell_step(lstm_cell_t* L, const float* x, int fast)
{
  const int I = L->input_size, H = L->hidden, W = I + H;
  for (int i = 0; i < I; i++)
    L->xh[i] = x[i];
  for (int r = 0; r < 4 * H; r++)
  {
    const float* wr = L->w + (size_t)r * W;
    float acc = 0.0f;
    for (int j = 0; j < W; j++)
      acc += wr[j] * L->xh[j];
    L->ifgo[r] = acc + L->b[r];
  }
  const float *gi = L->ifgo, *gf = L->ifgo + H, *gg = L->ifgo + 2 * H, *go = L->ifgo + 3 * H;
  if (fast)
  {
    for (int i = 0; i < H; i++)
      L->c[i] = fast_sigmoid_f(gf[i]) * L->c[i] + fast_sigmoid_f(gi[i]) * fast_tanh_f(gg[i]);
    for (int i = 0; i < H; i++)
      L->xh[I + i] = fast_sigmoid_f(go[i]) * fast_tanh_f(L->c[i]);
  }
  else
  {
    for (int i = 0; i < H; i++)
      L->c[i] = sigmoid_f(gf[i]) * L->c[i] + sigmoid_f(gi[i]) * tanhf(gg[i]);
    for (int i = 0; i < H; i++)
      L->xh[I + i] = sigmoid_f(go[i]) * tanhf(L->c[i]);
  }
}

/* Core processing on float planar buffers. */
static void convnet_process(nam_oracle* o, const float* const* in, float* const* out, int n);

static void process_core(nam_oracle* o, const float* const* in, float* const* out, int n)
{
  if (o->arch == NAM_ORACLE_ARCH_CONVNET)
  {
    convnet_process(o, in, out, n);
    return;
  }
  if (o->arch == NAM_ORACLE_ARCH_WAVENET)
  {
    /* _set_condition_array (model.cpp:809-820) */
    for (int ch = 0; ch < o->in_ch; ch++)
      for (int j = 0; j < n; j++)
        o->cond_in[(size_t)j * o->cond_dim + ch] = in[ch][j];
    /* _process_condition (model.cpp:777-807) */
    const float* cond = o->cond_in;
    if (o->cond_dsp)
    {
      nam_oracle* cd = o->cond_dsp;
      float* tmp_in[64];
      float* tmp_out[64];
      float* inb = (float*)xcalloc((size_t)cd->in_ch * n, sizeof(float));
      float* outb = (float*)xcalloc((size_t)cd->out_ch * n, sizeof(float));
      for (int ch = 0; ch < cd->in_ch; ch++)
      {
        tmp_in[ch] = inb + (size_t)ch * n;
        for (int j = 0; j < n; j++)
          tmp_in[ch][j] = o->cond_in[(size_t)j * o->cond_dim + ch];
      }
      for (int ch = 0; ch < cd->out_ch; ch++)
        tmp_out[ch] = outb + (size_t)ch * n;
      process_core(cd, (const float* const*)tmp_in, tmp_out, n);
      for (int ch = 0; ch < cd->out_ch; ch++)
        for (int j = 0; j < n; j++)
          o->cond_out[(size_t)j * cd->out_ch + ch] = tmp_out[ch][j];
      free(inb);
      free(outb);
      cond = o->cond_out;
    }
    const float* layer_in = o->cond_in; /* first array consumes the raw input (model.cpp:836-839) */
    for (int a = 0; a < o->n_arrays; a++)
    {
      larray_t* A = &o->arrays[a];
      if (a == 0)
        memset(A->head_inputs, 0, sizeof(float) * (size_t)A->head_out_size * n);
      else
        memcpy(A->head_inputs, o->arrays[a - 1].head_out, sizeof(float) * (size_t)A->head_out_size * n);
      layer_in = larray_process(A, layer_in, cond, n, o->fast_tanh);
    }
    const larray_t* last = &o->arrays[o->n_arrays - 1];
    if (o->with_head)
    {
      /* model.cpp:854-883 + Head::process :87-103 */
      pshead_t* H = &o->ph;
      for (long p = 0; p < (long)H->in_ch * n; p++)
        H->work[p] = o->head_scale * last->head_out[p];
      float* cur = H->work;
      for (int i = 0; i < H->n; i++)
      {
        act_apply(&H->act, cur, (long)H->convs[i].in_ch * n, o->fast_tanh);
        conv1d_process(&H->convs[i], cur, H->outs[i], n);
        cur = H->outs[i];
      }
      for (int ch = 0; ch < o->out_ch; ch++)
        for (int s = 0; s < n; s++)
          out[ch][s] = cur[(size_t)s * o->out_ch + ch];
    }
    else
    {
      for (int ch = 0; ch < o->out_ch; ch++)
        for (int s = 0; s < n; s++)
          out[ch][s] = o->head_scale * last->head_out[(size_t)s * o->out_ch + ch];
    }
  }
  else if (o->arch == NAM_ORACLE_ARCH_LSTM)
  {
    float xin[64], y[64];
    const int H = o->lstm_hidden;
    for (int i = 0; i < n; i++)
    {
      for (int ch = 0; ch < o->in_ch; ch++)
        xin[ch] = in[ch][i];
      if (o->n_cells == 0)
      {
        /* lstm.cpp:141-151 */
        const int m = o->in_ch < o->out_ch ? o->in_ch : o->out_ch;
        for (int ch = 0; ch < m; ch++)
          y[ch] = xin[ch];
        for (int ch = m; ch < o->out_ch; ch++)
          y[ch] = 0.0f;
      }
      else
      {
        lstm_cell_step(&o->cells[0], xin, o->fast_tanh);
        for (int l = 1; l < o->n_cells; l++)
          lstm_cell_step(&o->cells[l], o->cells[l - 1].xh + o->cells[l - 1].input_size, o->fast_tanh);
        const lstm_cell_t* L = &o->cells[o->n_cells - 1];
        const float* h = L->xh + L->input_size;
        for (int ch = 0; ch < o->out_ch; ch++)
        {
          float acc = 0.0f;
          for (int j = 0; j < H; j++)
            acc += o->head_w[(size_t)ch * H + j] * h[j];
          y[ch] = acc + o->head_b[ch];
        }
      }
      for (int ch = 0; ch < o->out_ch; ch++)
        out[ch][i] = y[ch];
    }
  }
  else if (o->arch == NAM_ORACLE_ARCH_LINEAR)
  {
    /* linear.cpp:168-199: out[i] = bias + dot(w_reversed, hist[i-RF+1 .. i]) */
    const int rf = o->rf;
    const int m = o->in_ch < o->out_ch ? o->in_ch : o->out_ch;
    for (int ch = 0; ch < o->in_ch; ch++)
    {
      float* h = o->lin_hist + (size_t)ch * (rf + o->max_buf);
      for (int j = 0; j < n; j++)
        h[rf + j] = in[ch][j];
    }
    for (int ch = 0; ch < m; ch++)
    {
      const float* h = o->lin_hist + (size_t)ch * (rf + o->max_buf);
      for (int i = 0; i < n; i++)
      {
        const float* x = h + i + 1; /* window [i+1, i+rf] == times [t-rf+1, t] */
        float acc = 0.0f;
        for (int j = 0; j < rf; j++)
          acc += o->lin_w[j] * x[j];
        out[ch][i] = o->lin_b + acc;
      }
    }
    for (int ch = m; ch < o->out_ch; ch++)
      for (int i = 0; i < n; i++)
        out[ch][i] = 0.0f;
    for (int ch = 0; ch < o->in_ch; ch++)
    {
      float* h = o->lin_hist + (size_t)ch * (rf + o->max_buf);
      memmove(h, h + n, sizeof(float) * (size_t)rf);
    }
  }
}

/* ConvNet::process (convnet.cpp:204-272): blocks of Conv1D(kernel 2) -> BatchNorm affine (:39-46: multiply, then
 * add) -> activation (:66-88), then the head W x + b (:155-170) */
static void convnet_process(nam_oracle* o, const float* const* in, float* const* out, int n)
{
  const int C = o->cn_channels;
  float *cur = o->cn_buf[0], *nxt = o->cn_buf[1];
  for (int f = 0; f < n; f++)
    for (int ch = 0; ch < o->in_ch; ch++)
      cur[(size_t)f * o->in_ch + ch] = in[ch][f];
  for (int i = 0; i < o->cn_blocks; i++)
  {
    conv1d_process(&o->cn_convs[i], cur, nxt, n);
    if (o->cn_batchnorm)
    {
      const float *sc = o->cn_scale + (size_t)i * C, *lc = o->cn_loc + (size_t)i * C;
      for (int f = 0; f < n; f++)
        for (int j = 0; j < C; j++)
        {
          float v = nxt[(size_t)f * C + j] * sc[j];
          v = v + lc[j];
          nxt[(size_t)f * C + j] = v;
        }
    }
    act_apply(&o->cn_act, nxt, (long)C * n, o->fast_tanh);
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  for (int f = 0; f < n; f++)
    for (int ch = 0; ch < o->out_ch; ch++)
    {
      float acc = 0.0f;
      for (int j = 0; j < C; j++)
        acc += o->cn_head_w[(size_t)ch * C + j] * cur[(size_t)f * C + j];
      out[ch][f] = acc + o->cn_head_b[ch];
    }
}

void nam_oracle_process_f32(nam_oracle* o, const float* const* in, float* const* out, int n)
{
  process_core(o, in, out, n);
}

void nam_oracle_process_f64(nam_oracle* o, const double* const* in, double* const* out, int n)
{
  float* buf = (float*)xcalloc((size_t)(o->in_ch + o->out_ch) * n, sizeof(float));
  const float* ip[64];
  float* op[64];
  for (int ch = 0; ch < o->in_ch; ch++)
  {
    float* p = buf + (size_t)ch * n;
    for (int j = 0; j < n; j++)
      p[j] = (float)in[ch][j]; /* the cast at model.cpp:817 / lstm.cpp:111 */
    ip[ch] = p;
  }
  for (int ch = 0; ch < o->out_ch; ch++)
    op[ch] = buf + (size_t)(o->in_ch + ch) * n;
  process_core(o, ip, op, n);
  for (int ch = 0; ch < o->out_ch; ch++)
    for (int j = 0; j < n; j++)
      out[ch][j] = (double)op[ch][j];
  free(buf);
}

/* DSP::prewarm (dsp.cpp:67-101) */
static void prewarm(nam_oracle* o)
{
  if (o->max_buf == 0)
    set_max_buffer(o, 4096); /* NAM_DEFAULT_MAX_BUFFER_SIZE */
  const int ps = o->prewarm_samples;
  if (ps == 0)
    return;
  const int bs = o->max_buf > 1 ? o->max_buf : 1;
  float* zin = (float*)xcalloc((size_t)o->in_ch * bs, sizeof(float));
  float* zout = (float*)xcalloc((size_t)o->out_ch * bs, sizeof(float));
  const float* ip[64];
  float* op[64];
  for (int ch = 0; ch < o->in_ch; ch++)
    ip[ch] = zin + (size_t)ch * bs;
  for (int ch = 0; ch < o->out_ch; ch++)
    op[ch] = zout + (size_t)ch * bs;
  int done = 0;
  while (done < ps)
  {
    process_core(o, ip, op, bs);
    done += bs;
  }
  free(zin);
  free(zout);
}

void nam_oracle_reset(nam_oracle* o, double sample_rate, int max_buffer_size, int do_prewarm)
{
  (void)sample_rate;
  set_max_buffer(o, max_buffer_size);
  if (do_prewarm)
    prewarm(o);
}

void nam_oracle_run_mono_f32(nam_oracle* o, const float* in, float* out, long n_total, int block)
{
  long pos = 0;
  while (pos < n_total)
  {
    const int n = (int)((n_total - pos) < block ? (n_total - pos) : block);
    const float* ip[1] = {in + pos};
    float* op[1] = {out + pos};
    process_core(o, ip, op, n);
    pos += n;
  }
}

/* ---- deep clone (state included) -------------------------------------------------------- */
static nam_oracle* clone_model(const nam_oracle* s)
{
  nam_oracle* d = (nam_oracle*)xcalloc(1, sizeof(nam_oracle));
  *d = *s;
  d->arrays = NULL;
  d->cells = NULL;
  d->cond_dsp = NULL;
  d->cond_in = d->cond_out = NULL;
  d->head_w = d->head_b = d->lin_w = d->lin_hist = NULL;
  memset(&d->ph, 0, sizeof(d->ph));
  if (s->arch == NAM_ORACLE_ARCH_WAVENET)
  {
    d->arrays = (larray_t*)xcalloc((size_t)s->n_arrays, sizeof(larray_t));
    for (int a = 0; a < s->n_arrays; a++)
    {
      const larray_t* SA = &s->arrays[a];
      larray_t* DA = &d->arrays[a];
      *DA = *SA;
      conv1x1_clone(&DA->rechannel, &SA->rechannel);
      conv1d_clone(&DA->head_rechannel, &SA->head_rechannel);
      DA->layers = (layer_t*)xcalloc((size_t)SA->n_layers, sizeof(layer_t));
      for (int l = 0; l < SA->n_layers; l++)
        layer_clone(&DA->layers[l], &SA->layers[l]);
      DA->rech_out = DA->head_inputs = DA->head_out = NULL;
      if (SA->rech_out)
      {
        const int mb = s->max_buf;
        DA->rech_out = (float*)xcalloc((size_t)SA->p.channels * mb, sizeof(float));
        DA->head_inputs = (float*)xcalloc((size_t)SA->head_out_size * mb, sizeof(float));
        DA->head_out = (float*)xcalloc((size_t)SA->p.head_size * mb, sizeof(float));
      }
    }
    if (s->with_head)
    {
      d->ph = s->ph;
      d->ph.convs = (conv1d_t*)xcalloc((size_t)s->ph.n, sizeof(conv1d_t));
      d->ph.outs = (float**)xcalloc((size_t)s->ph.n, sizeof(float*));
      for (int i = 0; i < s->ph.n; i++)
      {
        conv1d_clone(&d->ph.convs[i], &s->ph.convs[i]);
        if (s->max_buf)
          d->ph.outs[i] = (float*)xcalloc((size_t)s->ph.convs[i].out_ch * s->max_buf, sizeof(float));
      }
      act_clone(&d->ph.act, &s->ph.act);
      d->ph.work = s->max_buf ? (float*)xcalloc((size_t)s->ph.in_ch * s->max_buf, sizeof(float)) : NULL;
    }
    if (s->cond_dsp)
      d->cond_dsp = clone_model(s->cond_dsp);
    if (s->cond_in)
    {
      const int cod = s->cond_dsp ? s->cond_dsp->out_ch : s->cond_dim;
      d->cond_in = (float*)xcalloc((size_t)s->cond_dim * s->max_buf, sizeof(float));
      d->cond_out = (float*)xcalloc((size_t)cod * s->max_buf, sizeof(float));
    }
  }
  else if (s->arch == NAM_ORACLE_ARCH_LSTM)
  {
    const int H = s->lstm_hidden;
    d->cells = (lstm_cell_t*)xcalloc((size_t)s->n_cells, sizeof(lstm_cell_t));
    for (int l = 0; l < s->n_cells; l++)
    {
      const lstm_cell_t* S = &s->cells[l];
      lstm_cell_t* D = &d->cells[l];
      *D = *S;
      const int I = S->input_size;
      D->w = (float*)xcalloc((size_t)4 * H * (I + H), sizeof(float));
      memcpy(D->w, S->w, sizeof(float) * (size_t)4 * H * (I + H));
      D->b = (float*)xcalloc((size_t)4 * H, sizeof(float));
      memcpy(D->b, S->b, sizeof(float) * (size_t)4 * H);
      D->xh = (float*)xcalloc((size_t)(I + H), sizeof(float));
      memcpy(D->xh, S->xh, sizeof(float) * (size_t)(I + H));
      D->c = (float*)xcalloc((size_t)H, sizeof(float));
      memcpy(D->c, S->c, sizeof(float) * (size_t)H);
      D->ifgo = (float*)xcalloc((size_t)4 * H, sizeof(float));
    }
    d->head_w = (float*)xcalloc((size_t)s->out_ch * H, sizeof(float));
    memcpy(d->head_w, s->head_w, sizeof(float) * (size_t)s->out_ch * H);
    d->head_b = (float*)xcalloc((size_t)s->out_ch, sizeof(float));
    memcpy(d->head_b, s->head_b, sizeof(float) * (size_t)s->out_ch);
  }
  else if (s->arch == NAM_ORACLE_ARCH_LINEAR)
  {
    d->lin_w = (float*)xcalloc((size_t)s->rf, sizeof(float));
    memcpy(d->lin_w, s->lin_w, sizeof(float) * (size_t)s->rf);
    if (s->lin_hist)
    {
      const size_t nh = (size_t)s->in_ch * (size_t)(s->rf + s->max_buf);
      d->lin_hist = (float*)xcalloc(nh, sizeof(float));
      memcpy(d->lin_hist, s->lin_hist, nh * sizeof(float));
    }
  }
  return d;
}

typedef struct
{
  const nam_oracle* proto;
  const float* in;
  float* out;
  int batch, block;
  long n_total;
  int* next; /* shared work counter, claimed with an atomic add */
} batch_job_t;

static void* batch_worker(void* arg)
{
  batch_job_t* j = (batch_job_t*)arg;
  for (;;)
  {
    const int b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->batch)
      break;
    nam_oracle* m = clone_model(j->proto);
    nam_oracle_run_mono_f32(m, j->in + (size_t)b * j->n_total, j->out + (size_t)b * j->n_total, j->n_total, j->block);
    nam_oracle_destroy(m);
  }
  return NULL;
}

int nam_oracle_run_batch_mono_f32(const nam_oracle* proto, const float* in, float* out, int batch, long n_total,
                                  int block, int threads)
{
  if (proto->in_ch != 1 || proto->out_ch != 1 || proto->arch == NAM_ORACLE_ARCH_CONVNET)
  {
    set_err("run_batch_mono: model is not mono (or a ConvNet: no clone support, use one instance per stream)");
    return -1;
  }
  if (block > proto->max_buf)
  {
    set_err("run_batch_mono: block (%d) exceeds max_buffer_size (%d)", block, proto->max_buf);
    return -1;
  }
  if (threads < 1)
    threads = 1;
  if (threads > 256)
    threads = 256;
  int next = 0;
  batch_job_t job = {proto, in, out, batch, block, n_total, &next};
  pthread_t tid[256];
  int started = 0;
  for (int t = 1; t < threads; t++)
    if (pthread_create(&tid[started], NULL, batch_worker, &job) == 0)
      started++;
  batch_worker(&job);
  for (int t = 0; t < started; t++)
    pthread_join(tid[t], NULL);
  return 0;
}

/* ---- persistent batch: `batch` independent DSP instances that keep their state across calls ---- */
struct nam_oracle_batch
{
  int batch;
  nam_oracle** models;
};

typedef struct
{
  struct nam_oracle_batch* B;
  const float* in;
  float* out;
  long n_total, in_stride, out_stride;
  int block;
  int* next;
} pbatch_job_t;

static void* pbatch_worker(void* arg)
{
  pbatch_job_t* j = (pbatch_job_t*)arg;
  for (;;)
  {
    const int b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->B->batch)
      break;
    nam_oracle_run_mono_f32(j->B->models[b], j->in + (size_t)b * j->in_stride, j->out + (size_t)b * j->out_stride,
                            j->n_total, j->block);
  }
  return NULL;
}

struct nam_oracle_batch* nam_oracle_batch_create(const nam_oracle* proto, int batch)
{
  if (proto->in_ch != 1 || proto->out_ch != 1 || batch < 1 || proto->arch == NAM_ORACLE_ARCH_CONVNET)
  {
    set_err("batch_create: model must be mono (not a ConvNet) and batch >= 1");
    return NULL;
  }
  struct nam_oracle_batch* B = (struct nam_oracle_batch*)xcalloc(1, sizeof(*B));
  B->batch = batch;
  B->models = (nam_oracle**)xcalloc((size_t)batch, sizeof(nam_oracle*));
  for (int b = 0; b < batch; b++)
    B->models[b] = clone_model(proto);
  return B;
}

int nam_oracle_batch_process(struct nam_oracle_batch* B, const float* in, float* out, long n_total, long in_stride,
                             long out_stride, int block, int threads)
{
  if (block > B->models[0]->max_buf)
  {
    set_err("batch_process: block (%d) exceeds max_buffer_size (%d)", block, B->models[0]->max_buf);
    return -1;
  }
  if (threads < 1)
    threads = 1;
  if (threads > 256)
    threads = 256;
  int next = 0;
  pbatch_job_t job = {B, in, out, n_total, in_stride, out_stride, block, &next};
  pthread_t tid[256];
  int started = 0;
  for (int t = 1; t < threads; t++)
    if (pthread_create(&tid[started], NULL, pbatch_worker, &job) == 0)
      started++;
  pbatch_worker(&job);
  for (int t = 0; t < started; t++)
    pthread_join(tid[t], NULL);
  return 0;
}

void nam_oracle_batch_destroy(struct nam_oracle_batch* B)
{
  if (!B)
    return;
  for (int b = 0; b < B->batch; b++)
    nam_oracle_destroy(B->models[b]);
  free(B->models);
  free(B);
}

/* ------------------------------------------------------------------------------------------
 * Module-level entry points
 * ---------------------------------------------------------------------------------------- */
int nam_oracle_conv1d(int in_ch, int out_ch, int kernel, int dilation, int bias, int groups, const float* weights,
                      int n_weights, const float* in, float* out, int n, int n_calls)
{
  conv1d_t m;
  if (conv1d_init(&m, in_ch, out_ch, kernel, bias, dilation, groups))
    return -1;
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.w = weights;
  c.n_w = n_weights;
  conv1d_set_weights(&m, &c);
  if (c.failed || c.i_w != n_weights)
  {
    set_err("Conv1D weight count mismatch: consumed %d of %d", c.i_w, n_weights);
    conv1d_free(&m);
    return -1;
  }
  conv1d_set_max_buffer(&m, n);
  for (int k = 0; k < n_calls; k++)
    conv1d_process(&m, in + (size_t)k * n * in_ch, out + (size_t)k * n * out_ch, n);
  conv1d_free(&m);
  return 0;
}

int nam_oracle_conv1x1(int in_ch, int out_ch, int bias, int groups, const float* weights, int n_weights,
                       const float* in, float* out, int n)
{
  conv1x1_t m;
  if (conv1x1_init(&m, in_ch, out_ch, bias, groups))
    return -1;
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.w = weights;
  c.n_w = n_weights;
  conv1x1_set_weights(&m, &c);
  if (c.failed || c.i_w != n_weights)
  {
    set_err("Conv1x1 weight count mismatch: consumed %d of %d", c.i_w, n_weights);
    conv1x1_free(&m);
    return -1;
  }
  conv1x1_process(&m, in, in_ch, out, n);
  conv1x1_free(&m);
  return 0;
}

int nam_oracle_film(int cond_dim, int input_dim, int shift, int groups, const float* weights, int n_weights,
                    const float* in, const float* cond, float* out, int n)
{
  film_t f;
  if (film_init(&f, 1, cond_dim, input_dim, shift, groups))
    return -1;
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.w = weights;
  c.n_w = n_weights;
  conv1x1_set_weights(&f.css, &c);
  if (c.failed || c.i_w != n_weights)
  {
    set_err("FiLM weight count mismatch: consumed %d of %d", c.i_w, n_weights);
    film_free(&f);
    return -1;
  }
  film_set_max_buffer(&f, n);
  film_process(&f, in, input_dim, cond, cond_dim, n);
  memcpy(out, f.out, sizeof(float) * (size_t)input_dim * n);
  film_free(&f);
  return 0;
}

static void act_from_params(act_t* a, int type, const float* params, int n_params)
{
  int32_t cfg[2] = {type, n_params};
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.cfg = cfg;
  c.n_cfg = 2;
  c.fp = params;
  c.n_fp = n_params;
  act_parse(a, &c);
}

int nam_oracle_activation(int type, const float* params, int n_params, int fast_tanh, float* data, int channels,
                          int n)
{
  act_t a;
  act_from_params(&a, type, params, n_params);
  act_apply(&a, data, (long)channels * n, fast_tanh);
  act_free(&a);
  return 0;
}

int nam_oracle_gating(int mode, int act_type, const float* act_params, int n_act_params, int sec_type,
                      const float* sec_params, int n_sec_params, int channels, const float* in, float* out, int n)
{
  if (channels <= 0 || channels > 256 || (mode != GATING_GATED && mode != GATING_BLENDED))
  {
    set_err("gating: bad arguments");
    return -1;
  }
  act_t a, s;
  act_from_params(&a, act_type, act_params, n_act_params);
  act_from_params(&s, sec_type, sec_params, n_sec_params);
  float* z = (float*)xcalloc((size_t)2 * channels * n, sizeof(float));
  memcpy(z, in, sizeof(float) * (size_t)2 * channels * n);
  gate_apply(&a, &s, mode, channels, z, 2 * channels, n, 0);
  for (int f = 0; f < n; f++)
    memcpy(out + (size_t)f * channels, z + (size_t)f * 2 * channels, sizeof(float) * (size_t)channels);
  free(z);
  act_free(&a);
  act_free(&s);
  return 0;
}

/* One WaveNet Layer in isolation (model.cpp:183-393), for the reference's layer-level pins
 * (tools/test/test_wavenet/test_layer.cpp).  cfg = array params (as in the model cfg: 15 ints +
 * 8 x 3 FiLM ints), then kernel, dilation, gating_mode, ACT, SECONDARY_ACT.  Zero history. */
int nam_oracle_layer(const int32_t* cfg, int n_cfg, const float* fparams, int n_fparams, const float* weights,
                     int n_weights, const float* in, const float* cond, float* out_next, float* out_head, int n,
                     int fast_tanh)
{
  cursor_t c;
  memset(&c, 0, sizeof(c));
  c.cfg = cfg;
  c.n_cfg = n_cfg;
  c.fp = fparams;
  c.n_fp = n_fparams;
  c.w = weights;
  c.n_w = n_weights;
  array_params_t p;
  if (parse_array_params(&p, &c))
    return -1;
  const int kernel = take_i(&c), dilation = take_i(&c), gating = take_i(&c);
  layer_t L;
  if (layer_init(&L, &p, kernel, dilation, gating, &c))
  {
    layer_free(&L);
    return -1;
  }
  if (c.failed)
  {
    set_err("layer config stream truncated");
    layer_free(&L);
    return -1;
  }
  layer_set_weights(&L, &c);
  if (c.failed || c.i_w != n_weights)
  {
    set_err("Layer weight count mismatch: consumed %d of %d", c.i_w, n_weights);
    layer_free(&L);
    return -1;
  }
  layer_set_max_buffer(&L, n);
  layer_process(&L, in, cond, n, fast_tanh);
  memcpy(out_next, L.out_next, sizeof(float) * (size_t)L.channels * n);
  memcpy(out_head, L.out_head, sizeof(float) * (size_t)L.head_rows * n);
  layer_free(&L);
  return 0;
}
