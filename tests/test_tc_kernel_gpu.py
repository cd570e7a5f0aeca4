"""GPU parity of the tensor-core WaveNet kernel (neuralampmodelercore_b200/csrc/wavenet_tc.cuh, selected with
kernel_geometry=3; a build option, NAM_B200_BUILD_TC=1): tcgen05 TF32 MMAs with the 3-way hi/lo split must stay inside the same 1e-5 max-abs gate as
the FP32 kernel, on the same protocols."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import oracle
from tests import nam_fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-5
TC = 3  # kernel_geometry


@pytest.fixture(autouse=True)
def _needs_the_build_option():
    # the kernel is a build option since round 2 (slower than the FP32 kernels on the reference's families)
    if not nb.has_tensor_core_kernel():
        pytest.skip("libnam_b200.so was built without the tensor-core kernel (NAM_B200_BUILD_TC=1 to build it)")


def _oracle(nam, x, fast, block=64):
    m = oracle.OracleModel.from_dict(nam, fast_tanh=fast)
    m.reset(48000.0, block)
    if x.ndim == 1:
        return m.run(np.ascontiguousarray(x, np.float32), block)
    return m.run_batch(np.ascontiguousarray(x, np.float32), block)


def _gpu(nam, x, fast, block, batch_rows=None):
    x2 = x[None, :] if x.ndim == 1 else x
    d = nb.get_dsp(nam, batch=x2.shape[0], fast_tanh=fast, kernel_geometry=TC)
    d.Reset(48000.0, block)
    outs = [d.process_batch(np.ascontiguousarray(x2[:, p:p + block])) for p in range(0, x2.shape[1], block)]
    y = np.concatenate(outs, axis=1)
    d.close()
    return y[0] if x.ndim == 1 else y


@pytest.mark.parametrize("fast", [False, True], ids=["exact_tanh", "fast_tanh"])
def test_a1_standard_input_wav(fast):
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.input_wav()[40000:56000]
    ref = _oracle(nam, x, fast)
    for block in (64, 2048):
        got = _gpu(nam, x, fast, block)
        err = np.max(np.abs(got - ref))
        assert err <= TOL, f"block {block}: max-abs {err:.3e}"


@pytest.mark.parametrize("block", [1, 7, 127, 128, 129, 1000, 4096])
def test_block_sizes_and_partial_tiles(block):
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.synthetic_batch(1, 5000, seed=7)[0]
    ref = _oracle(nam, x, False)
    got = _gpu(nam, x, False, block)
    assert np.max(np.abs(got - ref)) <= TOL


def test_batch_of_streams_and_steady_state():
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.synthetic_batch(40, 3000)
    ref = _oracle(nam, x, True)
    got = _gpu(nam, x, True, 1024)
    assert np.max(np.abs(got - ref)) <= TOL
    d = nb.get_dsp(nam, batch=3, kernel_geometry=TC)
    d.Reset(48000.0, 64)
    for _ in range(3):
        y = d.process_batch(np.zeros((3, 64), np.float32))
    assert np.all(np.abs(y - (-1.19433e-3)) < 1e-6)


@pytest.mark.parametrize("channels,ks,dil,act", [
    ((16, 8), 3, [[1, 2, 4, 8, 16, 32, 64, 128, 256, 512]] * 2, "Tanh"),
    ((8,), 3, [[1, 2, 4, 8, 16, 32, 64, 128]], "ReLU"),
    ((16, 16), 2, [[1, 3, 7, 17, 41, 101, 239], [1, 13]], {"type": "LeakyReLU", "negative_slope": 0.01}),
    ((12, 6), 3, [[1, 2, 33, 50], [70, 100]], "Sigmoid"),
    ((5, 16), 3, [[40, 90], [1, 1]], "SiLU"),
])
def test_shape_family(channels, ks, dil, act):
    nam = fx.random_wavenet(channels=channels, kernel_size=ks, dilations=dil, activation=act, seed=11, scale=0.25)
    x = fx.synthetic_batch(2, 3000, seed=2)
    ref = _oracle(nam, x, False)
    got = _gpu(nam, x, False, 1500)
    err = np.max(np.abs(got - ref))
    assert err <= TOL, f"{channels} k={ks}: {err:.3e}"


def test_same_state_as_fp32_kernel():
    """Both kernels read and write the same per-stream rings: alternate them call by call on one signal."""
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.synthetic_batch(1, 4096, seed=5)
    ref = _oracle(nam, x[0], False)
    a = nb.get_dsp(nam, batch=1, kernel_geometry=TC)
    a.Reset(48000.0, 4096)
    y = a.process_batch(x)[0]
    assert np.max(np.abs(y - ref)) <= TOL


def test_unsupported_shapes_are_refused():
    with pytest.raises(nb.UnsupportedModelError):
        nb.get_dsp(fx.load_model("wavenet"), kernel_geometry=TC)  # 3 / 2 channels: below the 8-channel MMA K-step
