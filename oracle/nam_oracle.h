/*
 * nam_oracle.h -- CPU restatement of NeuralAmpModelerCore's per-sample inference path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
 * product (neuralampmodelercore_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  (1) Module level: every numeric known-answer the reference's own tests
 * hold for this path is reproduced in tests/test_oracle_pins.py.  (2) Whole-model level: the
 * reference ships no golden outputs and its Eigen submodule is missing, but its UNMODIFIED
 * sources compile here against oracle/eigen_shim (make ref -> oracle/_ref/libnam_ref.so), and
 * tests/test_reference_build.py holds this restatement to that build on all example models, both
 * tanh regimes, the A2 fast path and container semantics (max difference 0 .. 3e-6).
 *
 * The model is described by three flat arrays produced by oracle/nam_config.py from the
 * .nam JSON (schema documented there): int32 cfg[], float fparams[] (activation
 * parameters), float weights[] (the .nam "weights" array, untouched, in file order).
 */
#ifndef NAM_ORACLE_H
#define NAM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nam_oracle nam_oracle;

enum { NAM_ORACLE_ARCH_WAVENET = 1, NAM_ORACLE_ARCH_LSTM = 2, NAM_ORACLE_ARCH_LINEAR = 3, NAM_ORACLE_ARCH_CONVNET = 4 };

/* activation type codes: order of nam::activations::ActivationType (NAM/activations.h:26-39) */
enum {
  NAM_ACT_TANH = 0, NAM_ACT_HARDTANH, NAM_ACT_FASTTANH, NAM_ACT_RELU, NAM_ACT_LEAKYRELU, NAM_ACT_PRELU,
  NAM_ACT_SIGMOID, NAM_ACT_SILU, NAM_ACT_HARDSWISH, NAM_ACT_LEAKYHARDTANH, NAM_ACT_SOFTSIGN,
  NAM_ACT_IDENTITY = 100
};

/* Build a model.  fast_tanh mirrors Activation::enable_fast_tanh() having been called
 * BEFORE loading (NAM/activations.cpp:168-177).  condition_dsp may be NULL; when given, the
 * new model takes ownership (NAM/wavenet/model.cpp:919-930).  Returns NULL on error, the
 * message is available from nam_oracle_last_error(). */
nam_oracle* nam_oracle_create(const int32_t* cfg, int n_cfg, const float* fparams, int n_fparams,
                              const float* weights, int n_weights, double expected_sample_rate,
                              int fast_tanh, nam_oracle* condition_dsp);
void nam_oracle_destroy(nam_oracle* o);
const char* nam_oracle_last_error(void);

int nam_oracle_in_channels(const nam_oracle* o);
int nam_oracle_out_channels(const nam_oracle* o);
int nam_oracle_prewarm_samples(const nam_oracle* o);
/* number of weights consumed from the weights array while building (for loader tests) */
int nam_oracle_weights_consumed(const nam_oracle* o);

/* DSP::Reset(sampleRate, maxBufferSize): zero every history, then (if prewarm != 0) feed
 * zeros in maxBufferSize blocks until >= prewarm samples were processed
 * (NAM/dsp.cpp:67-101,130-140). */
void nam_oracle_reset(nam_oracle* o, double sample_rate, int max_buffer_size, int prewarm);

/* DSP::process: planar in[ch][frame] -> out[ch][frame], n <= max_buffer_size.
 * float and double variants (NAM_SAMPLE is double unless NAM_SAMPLE_FLOAT, NAM/dsp.h:18-22). */
void nam_oracle_process_f32(nam_oracle* o, const float* const* in, float* const* out, int n);
void nam_oracle_process_f64(nam_oracle* o, const double* const* in, double* const* out, int n);

/* Convenience for mono models: run a whole signal through in blocks of `block` frames. */
void nam_oracle_run_mono_f32(nam_oracle* o, const float* in, float* out, long n_total, int block);

/* Batch helper used for the CPU baseline: `batch` independent mono streams, each a clone of
 * `proto` (state copied as is, i.e. prewarmed), contiguous in[batch][n_total], processed in
 * blocks of `block` frames on `threads` pthreads. Returns 0 on success. */
int nam_oracle_run_batch_mono_f32(const nam_oracle* proto, const float* in, float* out, int batch, long n_total,
                                  int block, int threads);

/* Persistent batch for CPU-baseline timing: `batch` independent instances (clones of `proto`, state
 * included) that keep their own state across calls -- the reference's "one nam::DSP object per stream".
 * process(): stream b reads in[b*in_stride ..+n_total) in `block`-frame process() calls, on `threads`
 * pthreads (streams are distributed dynamically). */
typedef struct nam_oracle_batch nam_oracle_batch;
nam_oracle_batch* nam_oracle_batch_create(const nam_oracle* proto, int batch);
int nam_oracle_batch_process(nam_oracle_batch* b, const float* in, float* out, long n_total, long in_stride,
                             long out_stride, int block, int threads);
void nam_oracle_batch_destroy(nam_oracle_batch* b);

/* ---- module-level entry points (used to pin the restatement against the reference's own
 * known-answer unit tests) ---- */

/* Conv1D (NAM/conv1d.cpp): weights in .nam order (out,in,k per group) then bias.
 * Processes `n_calls` successive calls of `n` frames each from a zeroed history. */
int nam_oracle_conv1d(int in_ch, int out_ch, int kernel, int dilation, int bias, int groups, const float* weights,
                      int n_weights, const float* in /* in_ch x (n*n_calls) col-major */,
                      float* out /* out_ch x (n*n_calls) */, int n, int n_calls);
/* Conv1x1 (NAM/dsp.cpp:363-398,436-836) */
int nam_oracle_conv1x1(int in_ch, int out_ch, int bias, int groups, const float* weights, int n_weights,
                       const float* in, float* out, int n);
/* FiLM (NAM/film.h:76-190) */
int nam_oracle_film(int cond_dim, int input_dim, int shift, int groups, const float* weights, int n_weights,
                    const float* in, const float* cond, float* out, int n);
/* one activation applied to a (channels x n) column-major matrix in place */
int nam_oracle_activation(int type, const float* params, int n_params, int fast_tanh, float* data, int channels,
                          int n);
/* GatingActivation / BlendingActivation (NAM/gating_activations.h:60-114,166-228):
 * in is (2*channels x n), out is (channels x n). mode 1 = gated, 2 = blended. */
int nam_oracle_gating(int mode, int act_type, const float* act_params, int n_act_params, int sec_type,
                      const float* sec_params, int n_sec_params, int channels, const float* in, float* out, int n);

/* One WaveNet Layer (NAM/wavenet/model.cpp:183-393) from a zero history.  cfg = the per-array ints of
 * the model cfg (15 + 8x3), then kernel, dilation, gating_mode, ACT, SECONDARY_ACT.
 * out_next is (channels x n), out_head is (bottleneck or head1x1.out_channels x n). */
int nam_oracle_layer(const int32_t* cfg, int n_cfg, const float* fparams, int n_fparams, const float* weights,
                     int n_weights, const float* in, const float* cond, float* out_next, float* out_head, int n,
                     int fast_tanh);

#ifdef __cplusplus
}
#endif
#endif
