"""GPU parity of multi-channel models: in_channels / out_channels != 1 for WaveNet
(NAM/wavenet/model.cpp:809-820 input -> condition, :888-909 per-channel output; through the general kernel) and LSTM
(NAM/lstm.cpp:103-125 input vector, :79-97,164-167 out_channels x hidden head; thread-per-stream kernel).  The
batched entry takes a stream's row as its channel planes back to back: [batch][channels][frames]; DSP::process takes
the reference's NAM_SAMPLE** channel arrays.  Same 1e-5 gate against the oracle."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import nam_config, oracle
from tests import nam_fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _randomise(nam, seed, scale=0.35, last=None):
    n = nam_config.expected_weight_count(nam)
    w = np.random.default_rng(seed).uniform(-scale, scale, size=n).astype(np.float32)
    if last is not None:
        w[-1] = last
    nam["weights"] = [float(v) for v in w]
    return nam


def _signal(batch, channels, n, seed):
    x = fx.synthetic_batch(batch * channels, n, seed=seed).reshape(batch, channels, n)
    return np.ascontiguousarray(x)


def _oracle(nam, x, block=64, fast=False):
    """x: (batch, in_channels, n) -> (batch, out_channels, n), one oracle instance per stream, `block`-frame calls."""
    outs = []
    for b in range(x.shape[0]):
        m = oracle.OracleModel.from_dict(nam, fast_tanh=fast)
        m.reset(48000.0, block)
        y = [np.atleast_2d(m.process(np.ascontiguousarray(x[b, :, p:p + block]))) for p in range(0, x.shape[2], block)]
        outs.append(np.concatenate(y, axis=1))
        m.close()
    return np.stack(outs)


def _gpu(nam, x, block, **kw):
    d = nb.get_dsp(nam, batch=x.shape[0], **kw)
    assert (d.in_channels, d.out_channels) == (x.shape[1], d.out_channels)
    d.Reset(48000.0, block)
    y = np.concatenate([d.process_batch(np.ascontiguousarray(x[:, :, p:p + block])) for p in range(0, x.shape[2], block)],
                       axis=2)
    d.close()
    return y


def _wavenet_2in_3out():
    a0 = {"input_size": 2, "condition_size": 2, "head_size": 4, "channels": 5, "kernel_size": 3, "dilations": [1, 2, 4, 9],
          "activation": "Tanh", "gated": False, "head_bias": False}
    a1 = {"input_size": 5, "condition_size": 2, "head_size": 3, "channels": 4, "kernel_size": 3, "dilations": [1, 6],
          "activation": "Tanh", "gated": True, "head_bias": True}
    return _randomise(fx.make_wavenet_nam([a0, a1], [], head_scale=0.5, in_channels=2), seed=4, last=0.5)


@pytest.mark.parametrize("fast", [False, True], ids=["exact_tanh", "fast_tanh"])
def test_wavenet_two_in_three_out(fast):
    nam = _wavenet_2in_3out()
    info = nb.inspect(nam)
    assert info["kernel"] == "generic" and (info["in_channels"], info["out_channels"]) == (2, 3)
    x = _signal(5, 2, 1500, seed=12)
    ref = _oracle(nam, x, 64, fast)
    for block in (64, 700):
        err = np.max(np.abs(_gpu(nam, x, block, fast_tanh=fast) - ref))
        assert err <= TOL, f"block {block}: {err:.3e}"


def test_wavenet_dsp_process_channel_arrays():
    """nam::DSP::process(NAM_SAMPLE** in, NAM_SAMPLE** out, n): double planar, stream 0."""
    nam = _wavenet_2in_3out()
    x = _signal(1, 2, 640, seed=3)
    ref = _oracle(nam, x, 64)[0]
    d = nb.get_dsp(nam)
    d.Reset(48000.0, 64)
    out = np.zeros((3, 640), np.float64)
    for p in range(0, 640, 64):
        ins = [np.ascontiguousarray(x[0, c, p:p + 64], np.float64) for c in range(2)]
        outs = [np.zeros(64, np.float64) for _ in range(3)]
        d.process(ins, outs, 64)
        for c in range(3):
            out[c, p:p + 64] = outs[c]
    d.close()
    assert np.max(np.abs(out - ref)) <= TOL


@pytest.mark.parametrize("ci,co,H,nl", [(2, 3, 5, 1), (3, 1, 8, 2), (1, 2, 4, 1)])
def test_lstm_multichannel(ci, co, H, nl):
    nam = {"version": "0.5.4", "architecture": "LSTM", "sample_rate": 48000.0,
           "config": {"in_channels": ci, "out_channels": co, "input_size": ci, "hidden_size": H, "num_layers": nl},
           "weights": []}
    nam = _randomise(nam, seed=ci * 10 + co)
    x = _signal(70, ci, 600, seed=7)  # more streams than one 64-thread block
    for fast in (False, True):
        ref = _oracle(nam, x[:6], 64, fast)
        got = _gpu(nam, x, 256, fast_tanh=fast)
        err = np.max(np.abs(got[:6] - ref))
        assert err <= TOL, f"fast={fast}: {err:.3e}"
    assert got.shape == (70, co, 600)


def test_wrong_shapes_are_refused():
    d = nb.get_dsp(_wavenet_2in_3out(), batch=2)
    d.Reset(48000.0, 32)
    with pytest.raises(TypeError):
        d.process_batch(np.zeros((2, 32), np.float32))  # a multi-channel model takes [batch, channels, frames]
    with pytest.raises(TypeError):
        d.process_batch(np.zeros((2, 1, 32), np.float32))
    d.close()
