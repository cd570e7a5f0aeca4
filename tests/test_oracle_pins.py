"""Pin the CPU oracle (oracle/nam_oracle.c) against every numeric known-answer the reference's own unit
tests hold for the hot path.  Each test cites the reference test it transcribes
(paths relative to the reference repository, tools/test/).  Values and tolerances are the reference's.

These run on CPU (`-m "not gpu"`); they are what makes the oracle trustworthy as the GPU parity checker,
given that the reference itself cannot be built in this image (Eigen missing) and pins no whole-model output.
"""
import numpy as np
import pytest

from oracle import nam_config, oracle
from tests import nam_fixtures as fx


# ---------------------------------------------------------------------------------------------------
# Conv1D  -- test_conv1d.cpp
# ---------------------------------------------------------------------------------------------------
def test_conv1d_process_basic():
    # test_conv1d.cpp:161-207: K=2, weights {1,2}: out = 1*x[t-1] + 2*x[t], zero history
    y = oracle.conv1d([[1, 2, 3, 4]], [1.0, 2.0], 1, 1, 2, dilation=1, bias=False)
    assert np.allclose(y[0], [2.0, 5.0, 8.0, 11.0], atol=0.01)


def test_conv1d_process_with_bias():
    # test_conv1d.cpp:209-248
    y = oracle.conv1d([[2, 3]], [1.0, 0.0, 5.0], 1, 1, 2, bias=True)
    assert np.allclose(y[0], [5.0, 7.0], atol=0.01)


def test_conv1d_process_multichannel():
    # test_conv1d.cpp:250-303: (3x2) kernel-size-1 weights, row-major (out, in)
    y = oracle.conv1d([[1, 3], [2, 4]], [1, 0, 0, 1, 1, 1], 2, 3, 1, bias=False)
    assert np.allclose(y[:, 0], [1.0, 2.0, 3.0], atol=0.01)
    assert np.allclose(y[:, 1], [3.0, 4.0, 7.0], atol=0.01)


def test_conv1d_process_dilation():
    # test_conv1d.cpp:305-345: dilation 2: out = x[t-2] + 2 x[t]
    y = oracle.conv1d([[1, 2, 3, 4]], [1.0, 2.0], 1, 1, 2, dilation=2, bias=False)
    assert np.allclose(y[0], [2.0, 4.0, 7.0, 10.0], atol=0.01)


def test_conv1d_multiple_calls_keep_history():
    # test_conv1d.cpp:347-389: 3 calls of [1,2] with maxBufferSize 2 (forces a ring rewind); last call -> [3,3]
    y = oracle.conv1d([[1, 2, 1, 2, 1, 2]], [1.0, 1.0], 1, 1, 2, bias=False, n_calls=3)
    assert np.allclose(y[0, 4:], [3.0, 3.0], atol=0.01)
    assert np.allclose(y[0, :2], [1.0, 3.0], atol=0.01)  # first call sees the zero history


def test_conv1d_grouped_basic():
    # test_conv1d.cpp:499-558: 2 groups, identity / doubling blocks
    w = [1, 0, 0, 1, 2, 0, 0, 2]
    y = oracle.conv1d([[1, 5], [2, 6], [3, 7], [4, 8]], w, 4, 4, 1, bias=False, groups=2)
    assert np.allclose(y[:, 0], [1, 2, 6, 8], atol=0.01)
    assert np.allclose(y[:, 1], [5, 6, 14, 16], atol=0.01)


def test_conv1d_grouped_with_bias():
    # test_conv1d.cpp:560-613
    w = [1, 0, 0, 1, 1, 0, 0, 1, 1, 2, 3, 4]
    y = oracle.conv1d([[10], [20], [30], [40]], w, 4, 4, 1, bias=True, groups=2)
    assert np.allclose(y[:, 0], [11, 22, 33, 44], atol=0.01)


def _closed_form_conv1d(in_ch, out_ch, K, do_bias, dilation, n):
    # test_conv1d.cpp:18-99 (weights, bias, input and the brute-force expectation, all closed form)
    weights = []
    W = np.zeros((K, out_ch, in_ch), dtype=np.float32)
    for o in range(out_ch):
        for i in range(in_ch):
            for k in range(K):
                v = np.float32(0.011) * np.float32(o + 1) + np.float32(0.007) * np.float32(i + 1) - np.float32(0.003) * np.float32(k + 1)
                W[k, o, i] = v
                weights.append(v)
    bias = np.array([np.float32(-0.05) + np.float32(0.019) * np.float32(o + 1) for o in range(out_ch)], dtype=np.float32)
    if do_bias:
        weights += list(bias)
    x = np.zeros((in_ch, n), dtype=np.float32)
    for f in range(n):
        for i in range(in_ch):
            x[i, f] = np.float32(0.21) * (i + 1) - np.float32(0.037) * (f + 1) + np.float32(0.004) * ((i + 1) * (f + 1))
    exp = np.zeros((out_ch, n), dtype=np.float64)
    for f in range(n):
        for o in range(out_ch):
            s = float(bias[o]) if do_bias else 0.0
            for k in range(K):
                src = f - dilation * (K - 1 - k)
                if src < 0:
                    continue
                s += float(np.dot(W[k, o].astype(np.float64), x[:, src].astype(np.float64)))
            exp[o, f] = s
    return weights, x, exp


@pytest.mark.parametrize("shape", [(4, 8, 6, True, 3, 23), (4, 1, 16, True, 1, 23)])
def test_conv1d_matches_brute_force(shape):
    # test_conv1d.cpp:939-949, tolerance 1e-4 (:13-16)
    in_ch, out_ch, K, do_bias, dil, n = shape
    w, x, exp = _closed_form_conv1d(*shape)
    y = oracle.conv1d(x, w, in_ch, out_ch, K, dilation=dil, bias=do_bias)
    assert np.max(np.abs(y - exp)) < 1e-4


def test_conv1d_weight_count_checked():
    with pytest.raises(oracle.OracleError):
        oracle.conv1d([[1, 2]], [1.0], 1, 1, 2, bias=False)


# ---------------------------------------------------------------------------------------------------
# Conv1x1 -- test_conv_1x1.cpp
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(6, 4, False), (6, 8, False), (4, 4, True)])
def test_conv1x1_matches_reference_closed_form(shape):
    # test_conv_1x1.cpp:19-77,560-574, tolerance 1e-5 (:14-17)
    in_ch, out_ch, do_bias = shape
    n = 5
    W = np.zeros((out_ch, in_ch), dtype=np.float32)
    weights = []
    for o in range(out_ch):
        for i in range(in_ch):
            W[o, i] = np.float32(0.17) * (o + 1) - np.float32(0.031) * (i + 1)
            weights.append(W[o, i])
    b = np.array([np.float32(-0.09) + np.float32(0.023) * (o + 1) for o in range(out_ch)], dtype=np.float32)
    if do_bias:
        weights += list(b)
    x = np.zeros((in_ch, n), dtype=np.float32)
    for f in range(n):
        for i in range(in_ch):
            x[i, f] = np.float32(0.25) * (i + 1) - np.float32(0.11) * (f + 1) + np.float32(0.013) * ((i + 1) * (f + 1))
    exp = W.astype(np.float64) @ x.astype(np.float64) + (b[:, None] if do_bias else 0.0)
    y = oracle.conv1x1(x, weights, in_ch, out_ch, bias=do_bias)
    assert np.max(np.abs(y - exp)) < 1e-5


def test_conv1x1_grouped_block_diagonal():
    # test_conv_1x1.cpp grouped tests (:200-330): weights are per group (out_per_group x in_per_group)
    w = [1, 0, 0, 1, 2, 0, 0, 2]
    y = oracle.conv1x1([[1], [2], [3], [4]], w, 4, 4, bias=False, groups=2)
    assert np.allclose(y[:, 0], [1, 2, 6, 8])


def test_conv1x1_group_validation():
    # test_conv_1x1.cpp:102-137: channels must divide by groups
    with pytest.raises(oracle.OracleError):
        oracle.conv1x1([[0]] * 5, [0.0] * 30, 5, 6, bias=False, groups=2)
    with pytest.raises(oracle.OracleError):
        oracle.conv1x1([[0]] * 4, [0.0] * 20, 4, 5, bias=False, groups=2)


# ---------------------------------------------------------------------------------------------------
# Activations -- test_activations.cpp, NAM/activations.h:59-133
# ---------------------------------------------------------------------------------------------------
def test_fast_tanh_snapshot():
    # test_activations.cpp:17-31: fast_tanh(0) == 0 exactly
    assert oracle.activation(np.array([0.0], np.float32), "Fasttanh")[0] == 0.0


def test_leaky_relu_snapshots():
    # test_activations.cpp:69-79: (0,1,-1) -> (0,1,-0.01); :185-194 different slope
    y = oracle.activation(np.array([0.0, 1.0, -1.0], np.float32), "LeakyReLU")
    assert list(y) == [0.0, 1.0, np.float32(-0.01)]
    y = oracle.activation(np.array([-1.0], np.float32), {"type": "LeakyReLU", "negative_slope": 0.05})
    assert y[0] == np.float32(-0.05)


def test_softsign_snapshots():
    # test_activations.cpp:124-137
    y = oracle.activation(np.array([0.0, 1.0, -1.0, 2.0, -2.0], np.float32), "Softsign")
    assert np.allclose(y, [0.0, 0.5, -0.5, 2.0 / 3.0, -2.0 / 3.0], atol=1e-6)


def test_activation_formulas():
    # NAM/activations.h:59-133, one line each
    x = np.linspace(-4, 4, 41).astype(np.float32)
    x64 = x.astype(np.float64)
    sig = 1.0 / (1.0 + np.exp(-x64))
    cases = {
        "Tanh": np.tanh(x64),
        "Hardtanh": np.clip(x64, -1, 1),
        "ReLU": np.maximum(x64, 0),
        "Sigmoid": sig,
        "SiLU": x64 * sig,
        "Hardswish": x64 * np.clip(x64 + 3, 0, 6) / 6.0,
        "Softsign": x64 / (1 + np.abs(x64)),
    }
    for name, exp in cases.items():
        y = oracle.activation(x, name)
        assert np.max(np.abs(y - exp)) < 2e-6, name
    ax = np.abs(x64)
    ft = (x64 * (2.45550750702956 + 2.45550750702956 * ax + (0.893229853513558 + 0.821226666969744 * ax) * x64 * x64)
          / (2.44506634652299 + (2.44506634652299 + x64 * x64) * np.abs(x64 + 0.814642734961073 * x64 * ax)))
    assert np.max(np.abs(oracle.activation(x, "Fasttanh") - ft)) < 2e-6
    # enable_fast_tanh() swaps Tanh for Fasttanh (activations.cpp:168-177)
    assert np.array_equal(oracle.activation(x, "Tanh", fast_tanh=True), oracle.activation(x, "Fasttanh"))


def test_leaky_hardtanh_and_prelu():
    # activations.h:75-89 and :281-301 (slope picked by channel = pos % channels, column-major)
    x = np.array([[-2.0, 0.5, 3.0]], np.float32)
    y = oracle.activation(x, {"type": "LeakyHardtanh", "min_val": 0.0, "max_val": 0.9, "min_slope": 0.0, "max_slope": 0.02})
    assert np.allclose(y, [[0.0, 0.5, 0.9 + 0.02 * 2.1]], atol=1e-6)
    x2 = np.array([[-1.0, 2.0], [-1.0, -3.0]], np.float32)
    y2 = oracle.activation(x2, {"type": "PReLU", "negative_slopes": [0.04, 0.05]})
    assert np.allclose(y2, [[-0.04, 2.0], [-0.05, -0.15]], atol=1e-7)


# ---------------------------------------------------------------------------------------------------
# Gating / blending -- test_gating_activations.cpp, test_blending_detailed.cpp
# ---------------------------------------------------------------------------------------------------
def test_blending_identity_is_passthrough():
    # test_gating_activations.cpp:107-146: identity input activation -> output == input, any alpha
    x = np.array([[1.0, -1.0], [0.5, 0.8]], np.float32)
    for sec in (None, "Sigmoid"):
        y = oracle.gating(x, "blended", None, sec, 1)
        assert np.allclose(y, [[1.0, -1.0]], atol=1e-6)


def test_gating_product():
    # gating_activations.h:100-113: out = act(top) * act2(bottom)
    x = np.array([[0.5, -2.0], [0.0, 1.0]], np.float32)
    y = oracle.gating(x, "gated", "ReLU", "Sigmoid", 1)
    assert np.allclose(y, [[0.25, 0.0]], atol=1e-7)


def test_blending_formula():
    # gating_activations.h:209-227: alpha*act(x) + (1-alpha)*x
    x = np.array([[2.0], [0.0]], np.float32)
    y = oracle.gating(x, "blended", "Tanh", "Sigmoid", 1)
    assert abs(y[0, 0] - (0.5 * np.tanh(2.0) + 0.5 * 2.0)) < 1e-6


# ---------------------------------------------------------------------------------------------------
# FiLM -- test_film.cpp
# ---------------------------------------------------------------------------------------------------
def test_film_bias_only_scale_shift():
    # test_film.cpp:26-90: zero weights, biases = [scale(3), shift(3)] -> out = in*scale + shift (1e-6)
    cond_dim, dim = 2, 3
    w = [0.0] * (2 * dim * cond_dim) + [2.0, -1.0, 0.5, 10.0, -20.0, 3.0]
    x = np.array([[1, 2, 3, 4], [-1, -2, -3, -4], [0.25, 0.5, 0.75, 1.0]], np.float32)
    cond = np.random.default_rng(0).standard_normal((cond_dim, 4)).astype(np.float32)
    y = oracle.film(x, cond, w, cond_dim, dim, shift=True)
    exp = x * np.array([[2.0], [-1.0], [0.5]]) + np.array([[10.0], [-20.0], [3.0]])
    assert np.max(np.abs(y - exp)) < 1e-6


def test_film_scale_only():
    # test_film.cpp:92-140
    cond_dim, dim = 2, 3
    w = [0.0] * (dim * cond_dim) + [2.0, -1.0, 0.5]
    x = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    y = oracle.film(x, np.zeros((cond_dim, 2), np.float32), w, cond_dim, dim, shift=False)
    assert np.max(np.abs(y - x * np.array([[2.0], [-1.0], [0.5]]))) < 1e-6


def test_film_condition_dependent():
    # film.h:76-190 with non-zero weights: scale = W c + b
    w = [1.0, 0.5, 0.0, 1.0]  # cond 1 -> [scale, shift], weights (2x1) then bias (2)
    y = oracle.film([[2.0, 2.0]], [[1.0, 3.0]], w, 1, 1, shift=True)
    # scale = 1*c + 0, shift = 0.5*c + 1
    assert np.allclose(y, [[2.0 * 1.0 + 1.5, 2.0 * 3.0 + 2.5]], atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# WaveNet Layer -- test_wavenet/test_layer.cpp
# ---------------------------------------------------------------------------------------------------
def test_layer_gated_exact_values():
    # test_layer.cpp:39-116 ("Issue 101"): exact 0.5 / 0.25
    w = [1.0, 1.0, 0.0, 0.0, 1.0, -1.0, 1.0, 0.0]
    x = np.full((1, 4), 0.25, np.float32)
    nxt, head = oracle.layer(x, x, w, activation="ReLU", gating_mode="gated", secondary_activation="Sigmoid")
    assert np.all(nxt == 0.5)
    assert np.all(head == 0.25)


def test_layer_non_gated_values():
    # test_layer.cpp:143-209: layer output 3.0, head output 2.0
    w = [1.0, 0.0, 1.0, 1.0, 0.0]
    x = np.ones((1, 4), np.float32)
    nxt, head = oracle.layer(x, x, w, activation="ReLU", gating_mode="none")
    assert np.allclose(nxt, 3.0, atol=0.01)
    assert np.allclose(head, 2.0, atol=0.01)


def test_layer_weight_count_is_checked():
    with pytest.raises(oracle.OracleError):
        oracle.layer(np.ones((1, 2), np.float32), np.ones((1, 2), np.float32), [1.0, 0.0, 1.0], activation="ReLU")


# ---------------------------------------------------------------------------------------------------
# Whole WaveNet -- test_wavenet/test_full.cpp, test_wavenet/test_output_head.cpp
# ---------------------------------------------------------------------------------------------------
def _one_by_one_wavenet(head=None, extra_weights=(), head_scale=1.0):
    layers = [{"input_size": 1, "condition_size": 1, "head_size": 1, "channels": 1, "kernel_size": 1,
               "dilations": [1], "activation": "ReLU", "gated": False, "head_bias": False}]
    weights = [1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0, *extra_weights, head_scale]
    nam = fx.make_wavenet_nam(layers, weights, head_scale=head_scale)
    if head is not None:
        nam["config"]["head"] = head
    return nam


def test_wavenet_single_layer_analytic():
    # test_full.cpp:45-97 (weights given there; result follows from them): x=1 -> z=2 -> head 2 -> y = 2*head_scale
    m = oracle.OracleModel.from_dict(_one_by_one_wavenet(head_scale=0.5))
    m.reset(48000.0, 64)
    y = m.process(np.ones(4, np.float32))
    assert np.allclose(y, 1.0, atol=1e-6)


def test_wavenet_two_layer_post_stack_head():
    # test_output_head.cpp:117-186: |y| < 1e-6 for x = -0.25
    head = {"in_channels": 1, "channels": 1, "out_channels": 1, "kernel_sizes": [1, 1], "activation": "ReLU"}
    m = oracle.OracleModel.from_dict(_one_by_one_wavenet(head=head, extra_weights=[-1.0, 0.0, 2.0, 0.0]))
    m.reset(48000.0, 64)
    y = m.process(np.full(8, -0.25, np.float32))
    assert np.all(np.abs(y) < 1e-6)


def test_post_stack_head_receptive_field():
    # test_output_head.cpp:42-53: kernel sizes {3,5} -> head receptive field 7 -> prewarm 1 + 0 + 6
    head = {"channels": 1, "out_channels": 1, "kernel_sizes": [3, 5], "activation": "ReLU"}
    nam = _one_by_one_wavenet(head=head)
    nam["weights"] = [1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0] + [0.1] * (3 + 1 + 5 + 1) + [1.0]
    m = oracle.OracleModel.from_dict(nam)
    assert m.prewarm_samples == 1 + 0 + (7 - 1)


def test_head_scale_comes_from_last_weight():
    # NAM/wavenet/model.cpp:670: the JSON head_scale is ignored
    nam = _one_by_one_wavenet(head_scale=0.5)
    nam["config"]["head_scale"] = 123.0
    m = oracle.OracleModel.from_dict(nam)
    m.reset(48000.0, 8)
    assert np.allclose(m.process(np.ones(4, np.float32)), 1.0, atol=1e-6)


def test_weight_mismatch_errors():
    # NAM/wavenet/model.cpp:671-682
    nam = _one_by_one_wavenet()
    nam["weights"] = nam["weights"] + [0.0]
    with pytest.raises(oracle.OracleError, match="Weight mismatch"):
        oracle.OracleModel.from_dict(nam)
    nam["weights"] = nam["weights"][:-3]
    with pytest.raises(oracle.OracleError, match="expects more"):
        oracle.OracleModel.from_dict(nam)


# ---------------------------------------------------------------------------------------------------
# Linear -- test_linear.cpp
# ---------------------------------------------------------------------------------------------------
def test_linear_direct_known_values():
    # test_linear.cpp:100-111
    nam = {"version": "0.5.4", "architecture": "Linear", "config": {"receptive_field": 3, "bias": False},
           "weights": [0.5, -0.25, 0.125], "sample_rate": 48000}
    m = oracle.OracleModel.from_dict(nam)
    m.reset(48000.0, 4)
    y = m.process(np.array([1.0, 2.0, 3.0, 4.0], np.float64))
    assert np.allclose(y, [0.5, 0.75, 1.125, 1.5], atol=1e-7)


def test_linear_chunking_invariance():
    # test_linear.cpp:113-134 protocol (irregular chunks), direct form only
    rng = np.random.default_rng(3)
    rf = 37
    nam = {"version": "0.5.4", "architecture": "Linear", "config": {"receptive_field": rf, "bias": True},
           "weights": list(rng.uniform(-0.2, 0.2, rf + 1)), "sample_rate": 48000}
    x = rng.standard_normal(900).astype(np.float32)
    a = oracle.OracleModel.from_dict(nam)
    a.reset(48000.0, 900)
    ya = a.process(x)
    b = oracle.OracleModel.from_dict(nam)
    b.reset(48000.0, 512)
    yb, pos = [], 0
    for c in [1, 17, 64, 255, 3, 512, 31, 17]:
        yb.append(b.process(x[pos:pos + c]))
        pos += c
    assert pos == 900
    assert np.max(np.abs(np.concatenate(yb) - ya)) < 1e-6
    # brute force
    w = np.asarray(nam["weights"][:rf], np.float64)
    xp = np.concatenate([np.zeros(rf - 1), x.astype(np.float64)])
    exp = np.array([nam["weights"][rf] + np.dot(w, xp[t:t + rf][::-1]) for t in range(900)])
    assert np.max(np.abs(ya - exp)) < 1e-5


# ---------------------------------------------------------------------------------------------------
# LSTM -- the reference has NO numeric pin (test_lstm.cpp only checks isfinite); pin against a float64
# restatement of the equations in NAM/lstm.cpp:31-68,136-168 written independently here.
# ---------------------------------------------------------------------------------------------------
def _lstm_numpy(nam, x, fast=False):
    c = nam["config"]
    H, I, nl = c["hidden_size"], c["input_size"], c["num_layers"]
    w = np.asarray(nam["weights"], np.float64)
    pos = 0
    cells = []
    for l in range(nl):
        i_sz = I if l == 0 else H
        W = w[pos:pos + 4 * H * (i_sz + H)].reshape(4 * H, i_sz + H); pos += W.size
        b = w[pos:pos + 4 * H]; pos += 4 * H
        h = w[pos:pos + H].copy(); pos += H
        cc = w[pos:pos + H].copy(); pos += H
        cells.append([W, b, h, cc])
    hw = w[pos:pos + H]; pos += H
    hb = w[pos]; pos += 1
    assert pos == len(w)
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))  # noqa: E731
    out = []
    for xv in x:
        inp = np.array([xv], np.float64)
        for cell in cells:
            W, b, h, cc = cell
            q = W @ np.concatenate([inp, h]) + b
            i_, f_, g_, o_ = q[:H], q[H:2 * H], q[2 * H:3 * H], q[3 * H:]
            cc[:] = sig(f_) * cc + sig(i_) * np.tanh(g_)
            h[:] = sig(o_) * np.tanh(cc)
            inp = h
        out.append(hw @ cells[-1][2] + hb)
    return np.asarray(out)


def test_lstm_matches_float64_equations():
    nam = fx.load_model("lstm")
    x = fx.input_wav()[47000:49000]
    m = oracle.OracleModel.from_dict(nam)
    m.reset(48000.0, 64, prewarm=False)
    y = m.run(x, 64)
    exp = _lstm_numpy(nam, x.astype(np.float64))
    assert np.max(np.abs(y - exp)) < 2e-6


def test_lstm_prewarm_rounds_up_to_whole_blocks():
    # NAM/lstm.cpp:127-134 + NAM/dsp.cpp:95-100: 24000 samples at 48 kHz, fed in maxBufferSize blocks
    nam = fx.load_model("lstm")
    m = oracle.OracleModel.from_dict(nam)
    assert m.prewarm_samples == 24000
    rng = np.random.default_rng(0)
    # a model whose state has NOT converged distinguishes 24000 from 24576 steps: scale recurrent weights up
    nam2 = dict(nam)
    m1 = oracle.OracleModel.from_dict(nam2)
    m1.reset(48000.0, 1024)  # ceil(24000/1024)*1024 = 24576 zero steps
    m2 = oracle.OracleModel.from_dict(nam2)
    m2.reset(48000.0, 1024, prewarm=False)
    m2.run(np.zeros(24576, np.float32), 1024)
    x = rng.standard_normal(64).astype(np.float32) * 0.1
    assert np.array_equal(m1.process(x), m2.process(x))


# ---------------------------------------------------------------------------------------------------
# Loader: weight stream order / counts, prewarm counts
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,count,prewarm", [
    ("wavenet", 131, 23), ("wavenet_a1_standard", 13802, 4093), ("lstm", 70, 24000),
    ("a2_lite", 1871, 6347), ("a2_full", 12146, 6347), ("wavenet_a2_max", 818, None),
    ("wavenet_condition_dsp", 147, None),
])
def test_example_models_load_and_consume_every_weight(name, count, prewarm):
    nam = fx.load_model(name)
    assert len(nam["weights"]) == count
    assert nam_config.expected_weight_count(nam) == count
    m = oracle.OracleModel.from_dict(nam)
    assert m.weights_consumed == count
    if prewarm is not None:
        assert m.prewarm_samples == prewarm
    # tools/test/test_get_dsp.cpp:239-261 protocol: Reset, 3 x 64 frames, finite outputs
    m.reset(48000.0, 64)
    for _ in range(3):
        y = m.process(np.zeros((m.in_channels, 64), np.float32))
        assert np.all(np.isfinite(y))


def test_prewarmed_state_is_steady():
    # fact 6 of SURVEY.md: after Reset the history is the non-zero steady state; zeros in -> constant out
    m = oracle.OracleModel.from_dict(fx.load_model("wavenet_a1_standard"))
    m.reset(48000.0, 64)
    y = m.process(np.zeros(64, np.float32))
    assert np.all(y == y[0]) and abs(y[0] - (-1.19433e-3)) < 1e-7  # SURVEY.md Appendix E probe value
    m2 = oracle.OracleModel.from_dict(fx.load_model("wavenet"))
    m2.reset(48000.0, 64)
    assert abs(m2.process(np.zeros(8, np.float32))[0] - (-4.8155e-4)) < 1e-7


@pytest.mark.parametrize("name", ["wavenet", "wavenet_a1_standard", "lstm"])
@pytest.mark.parametrize("regime", ["exact", "fast"])
def test_oracle_matches_committed_golden(name, regime):
    """Regression pin of the oracle itself on example_audio/input.wav (render.cpp protocol: Reset(sr,64),
    64-frame blocks).  tests/golden/make_golden.py produced the vectors with the strict build."""
    gold = fx.oracle_golden(name, regime)
    m = oracle.OracleModel.from_dict(fx.load_model(name), fast_tanh=(regime == "fast"))
    m.reset(48000.0, 64)
    x = fx.input_wav()
    n = 50096 if name == "wavenet_a1_standard" else len(x)  # keep the CPU suite short
    y = m.run(x[:n], 64)
    assert np.max(np.abs(y[:512] - gold["head"])) < 1e-7
    assert np.max(np.abs(y[46000:50096] - gold["transition"])) < 1e-7
    if n == len(x):
        assert np.max(np.abs(y[::37] - gold["strided"])) < 1e-7


def test_block_size_invariance():
    # the path is causal and stateful: output must not depend on how the stream is chunked
    nam = fx.load_model("wavenet")
    x = fx.input_wav()[47000:50000]
    ref = None
    for block in (1, 7, 64, 1000, 3000):
        m = oracle.OracleModel.from_dict(nam)
        m.reset(48000.0, block)
        y = m.run(x, block)
        if ref is None:
            ref = y
        assert np.max(np.abs(y - ref)) < 1e-7


def test_double_boundary_casts_to_float():
    # NAM/wavenet/model.cpp:809-820,888-897: NAM_SAMPLE=double is cast to float at the edge
    nam = fx.load_model("wavenet")
    x = (fx.input_wav()[48000:48256]).astype(np.float64) + 1e-12
    a = oracle.OracleModel.from_dict(nam)
    a.reset(48000.0, 256)
    ya = a.process(x)
    b = oracle.OracleModel.from_dict(nam)
    b.reset(48000.0, 256)
    yb = b.process(x.astype(np.float32))
    assert ya.dtype == np.float64 and np.array_equal(ya.astype(np.float32), yb)


def test_version_gate():
    # NAM/get_dsp.cpp:18-39
    assert nam_config.version_support("0.5.0") == "yes"
    assert nam_config.version_support("0.7.0") == "yes"
    assert nam_config.version_support("0.7.3") == "partial"
    assert nam_config.version_support("0.4.9") == "no"
    assert nam_config.version_support("0.8.0") == "no"
    assert nam_config.version_support("1.0.0") == "no"
    assert nam_config.version_support("0.5") == "no"
