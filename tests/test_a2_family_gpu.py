"""GPU parity of the A2 model family: 23 layers, per-layer kernel sizes 6 / 15, LeakyReLU, odd
dilations up to 239 and a kernel-16 head convolution (NAM/wavenet/model.cpp:397-400,548; the shape the
reference special-cases in NAM/wavenet/a2_fast.cpp).  Same 1e-5 gate, same protocols as tests/test_parity_gpu.py;
the tolerance the reference accepts between its own two implementations of this shape is 5e-5
(tools/test/test_a2_fast.cpp:296-298)."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import oracle
from tests import nam_fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-5


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


@pytest.mark.parametrize("name", ["a2_lite", "a2_full"])
def test_input_wav(name):
    nam = fx.load_model(name)
    x = fx.input_wav()[44000:56000]  # silence -> sine transition
    ref = _oracle(nam, x)
    for block in (64, 4096):
        err = np.max(np.abs(_gpu(nam, x, block) - ref))
        assert err <= TOL, f"{name} block {block}: {err:.3e}"


@pytest.mark.parametrize("name", ["a2_lite", "a2_full"])
def test_matches_committed_golden(name):
    """The oracle's own outputs for the full 2 s input.wav are committed (tests/golden/oracle_outputs.npz)."""
    nam = fx.load_model(name)
    gold = fx.oracle_golden(name, "exact")
    got = _gpu(nam, fx.input_wav(), 4096)
    for key, val in fx.decimate(got).items():
        err = np.max(np.abs(val - gold[key]))
        assert err <= TOL, f"{name} {key}: {err:.3e}"


@pytest.mark.parametrize("block", [1, 15, 16, 17, 333, 512, 513])
def test_block_sizes_cross_the_head_history(block):
    nam = fx.load_model("a2_full")
    x = fx.synthetic_batch(1, 2600, seed=21)[0]
    ref = _oracle(nam, x)
    assert np.max(np.abs(_gpu(nam, x, block) - ref)) <= TOL


def test_batch_and_prewarmed_silence():
    nam = fx.load_model("a2_full")
    x = fx.synthetic_batch(37, 3000, seed=4)
    ref = _oracle(nam, x)
    assert np.max(np.abs(_gpu(nam, x, 1000) - ref)) <= TOL
    # benchmodel protocol: zeros in, the prewarmed steady state out
    z = np.zeros((2, 192), np.float32)
    assert np.max(np.abs(_gpu(nam, z, 64) - _oracle(nam, z))) <= 1e-6


@pytest.mark.parametrize("channels,hk,hd", [((3,), 16, None), ((8,), 16, None), ((16,), 5, 3), ((6, 4), 9, 2), ((16, 8), 33, 2)])
def test_head_convolution_shapes(channels, hk, hd):
    """Head Conv1D with kernel > 1 (and head_dilation) on every array, 1 and 2 arrays, look-back up to 64."""
    dil = [[1, 3, 7, 17, 41], [1, 13]][:len(channels)]
    nam = fx.random_wavenet(channels=channels, kernel_size=3, dilations=dil, head_kernel=hk, head_dilation=hd,
                            activation={"type": "LeakyReLU", "negative_slope": 0.01}, seed=3, scale=0.25)
    x = fx.synthetic_batch(3, 2500, seed=8)
    ref = _oracle(nam, x)
    for geom in (1, 2):
        err = np.max(np.abs(_gpu(nam, x, 700, kernel_geometry=geom) - ref))
        assert err <= TOL, f"{channels} hk={hk} geometry {geom}: {err:.3e}"


def test_head_lookback_beyond_the_tile_halo_goes_to_the_general_kernel():
    """The fused kernel keeps at most 64 frames of head history in its tile halo; longer head look-backs are served
    by the general kernel instead (same parity gate)."""
    nam = fx.random_wavenet(channels=(4,), dilations=[[1, 2]], head_kernel=16, head_dilation=5, seed=1)
    info = nb.inspect(nam)
    assert info["kernel"] == "generic" and "looks back 75" in info["reason"]
    x = fx.synthetic_batch(2, 900, seed=4)
    assert np.max(np.abs(_gpu(nam, x, 300) - _oracle(nam, x))) <= TOL


def test_slimmable_container_dispatch():
    """example_models/A2.nam is a SlimmableContainer of A2-Lite (max_value 0.5) and A2-Full (1.0): the full size is
    active by default, SetSlimmableSize switches, the newly active sub-model starts from its own prewarmed state
    (NAM/container.cpp:49,88-133)."""
    lite, full = fx.load_model("a2_lite"), fx.load_model("a2_full")
    nam = fx.make_container([(0.5, lite), (1.0, full)])
    x = fx.synthetic_batch(3, 1500, seed=12)
    d = nb.get_dsp(nam, batch=3)
    assert d.GetSlimmableSizeBreakpoints() == [0.5]
    d.Reset(48000.0, 500)
    y_full = np.concatenate([d.process_batch(np.ascontiguousarray(x[:, p:p + 500])) for p in (0, 500)], axis=1)
    assert np.max(np.abs(y_full - _oracle(full, x[:, :1000]))) <= TOL
    d.SetSlimmableSize(0.2)
    y_lite = np.concatenate([d.process_batch(np.ascontiguousarray(x[:, p:p + 500])) for p in (0, 500, 1000)], axis=1)
    assert np.max(np.abs(y_lite - _oracle(lite, x))) <= TOL
    d.SetSlimmableSize(0.5)  # 0.5 is not < 0.5: back to the full model, freshly reset
    y2 = d.process_batch(np.ascontiguousarray(x[:, :500]))
    assert np.max(np.abs(y2 - _oracle(full, x[:, :500]))) <= TOL
    d.SetSlimmableSize(0.99)  # already active: no reset, the stream continues
    y3 = d.process_batch(np.ascontiguousarray(x[:, 500:1000]))
    assert np.max(np.abs(y3 - _oracle(full, x[:, :1000])[:, 500:])) <= TOL
    d.close()
    plain = nb.get_dsp(full)
    with pytest.raises(nb.UnsupportedModelError):
        plain.SetSlimmableSize(0.3)
    assert plain.GetSlimmableSizeBreakpoints() == []
    plain.close()
