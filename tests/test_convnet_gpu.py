"""GPU parity of the ConvNet architecture (NAM/convnet.cpp): blocks of kernel-2 dilated Conv1D ->
BatchNorm -> activation and a linear head, mono and multi-channel, with and without batchnorm, grouped convolutions.
The reference's own tests assert only isfinite (tools/test/test_convnet.cpp); the oracle is pinned to the reference
build on the same configurations (tests/test_reference_build.py::test_convnet).  Same 1e-5 gate."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import oracle
from tests import nam_fixtures as fx
from tests.test_reference_build import CONVNETS, _convnet

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle(nam, x, block=64, fast=False):
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
    d.Reset(48000.0, block)
    mono = d.in_channels == 1 and d.out_channels == 1
    ys = []
    for p in range(0, x.shape[2], block):
        xb = np.ascontiguousarray(x[:, :, p:p + block])
        ys.append(d.process_batch(np.ascontiguousarray(xb[:, 0]))[:, None] if mono else d.process_batch(xb))
    d.close()
    return np.concatenate(ys, axis=2)


@pytest.mark.parametrize("kw", CONVNETS, ids=[f"convnet{i}" for i in range(len(CONVNETS))])
@pytest.mark.parametrize("fast", [False, True], ids=["exact_tanh", "fast_tanh"])
def test_convnet_matches_oracle(kw, fast):
    nam = _convnet(**kw)
    info = nb.inspect(nam)
    assert info["kernel"] == "convnet" and info["prewarm_samples"] == 1 + sum(kw["dilations"])
    ci = kw.get("in_channels", 1)
    x = fx.synthetic_batch(3 * ci, 1100, seed=31).reshape(3, ci, 1100)
    ref = _oracle(nam, x, 64, fast)
    for block in (64, 500):
        err = np.max(np.abs(_gpu(nam, x, block, fast_tanh=fast) - ref))
        assert err <= TOL, f"block {block}: {err:.3e}"


def test_convnet_standard_shape_batch():
    """The legacy 'standard' ConvNet shape (16 channels, dilations 1..1024 and a second short ladder, batchnorm, Tanh):
    70 streams (more than one 64-wide scheduling unit), prewarm from Reset, state carried across calls."""
    nam = _convnet(channels=16, dilations=[2 ** i for i in range(11)] + [1, 2, 4, 8], batchnorm=True, activation="Tanh", seed=9)
    x = fx.synthetic_batch(70, 900, seed=4).reshape(70, 1, 900)
    got = _gpu(nam, x, 300)
    ref = _oracle(nam, x[:4], 60)
    assert np.max(np.abs(got[:4] - ref)) <= TOL
    d = nb.get_dsp(nam, batch=2)
    d.Reset(48000.0, 64)
    z = d.process_batch(np.zeros((2, 64), np.float32))  # prewarmed with zeros: the steady state of a silent input
    d.Reset(48000.0, 64)
    assert np.array_equal(z, d.process_batch(np.zeros((2, 64), np.float32)))
    d.close()
