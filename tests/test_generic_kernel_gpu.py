"""GPU parity of the general WaveNet kernel (neuralampmodelercore_b200/csrc/wavenet_generic.cuh):
gated / blended activations, bottleneck != channels, grouped convolutions, head1x1, the eight FiLM sites, a
condition_dsp sub-model and the post-stack head (NAM/wavenet/model.cpp:19-103,183-393,777-910, NAM/film.h,
NAM/gating_activations.h) -- the options of example_models/wavenet_a2_max.nam and wavenet_condition_dsp.nam, which
the reference's own end-to-end smoke test loads (tools/test/test_get_dsp.cpp:243-245).  Same 1e-5 gate against the
oracle; kernel_geometry=4 also pushes the fused-family models through it."""
import json

import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import nam_config, oracle
from tests import nam_fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-5
GENERAL = 4  # kernel_geometry


def _oracle(nam, x, block=64, fast=False):
    m = oracle.OracleModel.from_dict(nam, fast_tanh=fast)
    m.reset(48000.0, block)
    x = np.ascontiguousarray(x, np.float32)
    return m.run(x, block) if x.ndim == 1 else m.run_batch(x, block)


def _gpu(nam, x, block, **kw):
    x2 = x[None, :] if x.ndim == 1 else x
    d = nb.get_dsp(nam, batch=x2.shape[0], **kw)
    d.Reset(48000.0, block)
    y = np.concatenate([d.process_batch(np.ascontiguousarray(x2[:, p:p + block])) for p in range(0, x2.shape[1], block)],
                       axis=1)
    d.close()
    return y[0] if x.ndim == 1 else y


def _randomise(nam, seed, scale=0.4, head_scale=0.5):
    n = nam_config.expected_weight_count(nam)
    w = np.random.default_rng(seed).uniform(-scale, scale, size=n).astype(np.float32)
    w[-1] = head_scale
    nam["weights"] = [float(v) for v in w]
    return nam


@pytest.mark.parametrize("name", ["wavenet_condition_dsp", "wavenet_a2_max"])
@pytest.mark.parametrize("fast", [False, True], ids=["exact_tanh", "fast_tanh"])
def test_example_models(name, fast):
    nam = fx.load_model(name)
    assert nb.inspect(nam)["kernel"] == "generic"
    x = fx.input_wav()[46000:52000]
    ref = _oracle(nam, x, 64, fast)
    for block in (64, 1000):
        err = np.max(np.abs(_gpu(nam, x, block, fast_tanh=fast) - ref))
        assert err <= TOL, f"{name} block {block}: {err:.3e}"
    xb = fx.synthetic_batch(70, 900, seed=3)  # more streams than one 64-thread block
    err = np.max(np.abs(_gpu(nam, xb, 300, fast_tanh=fast) - _oracle(nam, xb, 64, fast)))
    assert err <= TOL, f"{name} batch: {err:.3e}"


def test_prewarmed_silence_and_reset():
    nam = fx.load_model("wavenet_a2_max")
    z = np.zeros((3, 128), np.float32)
    ref = _oracle(nam, z)
    assert np.max(np.abs(_gpu(nam, z, 64) - ref)) <= 1e-6
    d = nb.get_dsp(nam, batch=3)
    d.Reset(48000.0, 64)
    x = fx.synthetic_batch(3, 64, seed=1)
    a = d.process_batch(x)
    d.Reset(48000.0, 64)  # Reset returns every stream to the prewarmed state
    b = d.process_batch(x)
    assert np.array_equal(a, b)
    d.close()


@pytest.mark.parametrize("name", ["wavenet", "wavenet_a1_standard", "a2_lite"])
def test_fused_family_models_through_the_general_kernel(name):
    nam = fx.load_model(name)
    x = fx.synthetic_batch(2, 700, seed=6)
    ref = _oracle(nam, x)
    err = np.max(np.abs(_gpu(nam, x, 256, kernel_geometry=GENERAL) - ref))
    assert err <= TOL, f"{name}: {err:.3e}"


def _film(active=True, shift=True, groups=1):
    return {"active": active, "shift": shift, "groups": groups}


def test_gating_modes_groups_head1x1_and_bottleneck():
    layer = {
        "input_size": 1, "condition_size": 1, "channels": 8, "bottleneck": 4,
        "head": {"out_channels": 1, "kernel_size": 3, "bias": True},
        "kernel_sizes": [3, 2, 4], "dilations": [1, 2, 5],
        "activation": ["Tanh", {"type": "LeakyReLU", "negative_slope": 0.02}, "ReLU"],
        "gating_mode": ["gated", "blended", "none"], "secondary_activation": ["Sigmoid", "Hardswish", "Sigmoid"],
        "groups_input": 2, "groups_input_mixin": 1,
        "layer1x1": {"active": True, "groups": 2}, "head1x1": {"active": True, "out_channels": 6, "groups": 2},
        "layer1x1_post_film": _film(groups=1), "activation_post_film": _film(shift=False),
    }
    nam = _randomise(fx.make_wavenet_nam([layer], [], version="0.7.0"), seed=21)
    assert nb.inspect(nam)["kernel"] == "generic"
    x = fx.synthetic_batch(4, 1200, seed=2)
    err = np.max(np.abs(_gpu(nam, x, 500) - _oracle(nam, x)))
    assert err <= TOL, f"{err:.3e}"


def test_two_arrays_without_layer1x1_and_post_stack_head():
    a0 = {"input_size": 1, "condition_size": 1, "channels": 6, "bottleneck": 6, "head_size": 4, "head_bias": False,
          "kernel_size": 3, "dilations": [1, 3], "activation": "Softsign", "gating_mode": "none",
          "layer1x1": {"active": False, "groups": 1},
          "conv_pre_film": _film(), "conv_post_film": _film(shift=False), "input_mixin_pre_film": _film(),
          "input_mixin_post_film": _film(), "activation_pre_film": _film()}
    a1 = {"input_size": 6, "condition_size": 1, "channels": 4, "bottleneck": 4, "head_size": 3, "head_bias": True,
          "kernel_size": 2, "dilations": [2, 7], "activation": {"type": "PReLU", "negative_slopes": [0.1, 0.2, 0.05, 0.3]},
          "gating_mode": "none"}
    head = {"in_channels": 3, "channels": 5, "out_channels": 1, "kernel_sizes": [3, 2], "activation": "ReLU"}
    nam = fx.make_wavenet_nam([a0, a1], [], version="0.7.0")
    nam["config"]["head"] = head
    nam = _randomise(nam, seed=8)
    info = nb.inspect(nam)
    assert info["kernel"] == "generic", info
    x = fx.synthetic_batch(3, 1500, seed=5)
    err = np.max(np.abs(_gpu(nam, x, 640) - _oracle(nam, x)))
    assert err <= TOL, f"{err:.3e}"


def test_too_wide_for_the_general_kernel_is_refused():
    layer = {"input_size": 1, "condition_size": 1, "channels": 40, "bottleneck": 40, "head_size": 1, "head_bias": True,
             "kernel_size": 2, "dilations": [1], "activation": "Tanh", "gating_mode": "gated"}
    nam = _randomise(fx.make_wavenet_nam([layer], [], version="0.7.0"), seed=1)
    info = nb.inspect(nam)
    assert info["kernel"] == "unsupported" and "general kernel" in info["reason"]
    with pytest.raises(nb.UnsupportedModelError):
        nb.get_dsp(nam)
