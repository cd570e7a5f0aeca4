"""GPU parity of the tile-parallel launch modes of the fused WaveNet kernel (few streams x long calls): the
lock-step mode (default: every (stream, tile) its own CTA, hand-over through the per-call history buffer), the
wavefront mode (tile_mode=1) and the serial walk (tile_mode=2) must all reproduce the oracle within the same
1e-5 gate, for both tile geometries, and may be interleaved freely on one handle (the rings are the only state
that survives a call)."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import oracle
from tests import nam_fixtures as fx

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle(nam, x, fast=False):
    m = oracle.OracleModel.from_dict(nam, fast_tanh=fast)
    m.reset(48000.0, 64)
    x = np.ascontiguousarray(x, np.float32)
    return m.run(x, 64) if x.ndim == 1 else m.run_batch(x, 64)


def _gpu_chunks(nam, x, chunks, max_frames, **kw):
    x2 = x[None, :] if x.ndim == 1 else x
    d = nb.get_dsp(nam, batch=x2.shape[0], **kw)
    d.Reset(48000.0, max_frames)
    out, pos, i = [], 0, 0
    while pos < x2.shape[1]:
        n = min(chunks[i % len(chunks)], x2.shape[1] - pos)
        out.append(d.process_batch(np.ascontiguousarray(x2[:, pos:pos + n])))
        pos += n
        i += 1
    d.close()
    y = np.concatenate(out, axis=1)
    return y[0] if x.ndim == 1 else y


@pytest.mark.parametrize("tile_mode", [0, 1, 2], ids=["lockstep", "wavefront", "serial"])
@pytest.mark.parametrize("geom", [0, 1, 2], ids=["auto", "tile256", "tile512"])
def test_a1_standard_long_and_short_calls_interleaved(tile_mode, geom):
    """Calls of 3000 / 5000 frames (tile-parallel) between calls of 700 / 64 / 1 frames (one tile, classic walk):
    the look-backs (up to 1024 frames) reach across call boundaries in every combination."""
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.synthetic_batch(3, 11000, seed=11)
    ref = _oracle(nam, x, fast=True)
    got = _gpu_chunks(nam, x, [3000, 700, 5000, 64, 1, 1235], 5000, fast_tanh=True, tile_mode=tile_mode, kernel_geometry=geom)
    err = np.max(np.abs(got - ref))
    assert err <= TOL, f"tile_mode {tile_mode} geometry {geom}: {err:.3e}"


@pytest.mark.parametrize("tile_mode", [0, 1])
def test_one_call_for_the_whole_file(tile_mode):
    """benchmodel's 2 s of audio (tools/benchmodel.cpp:105) as ONE 96,000-frame call of a single stream: 375 tiles of
    256 frames in flight at once in the lock-step mode."""
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.input_wav()
    gold = fx.oracle_golden("wavenet_a1_standard", "exact")
    y = _gpu_chunks(nam, x, [len(x)], len(x), fast_tanh=False, tile_mode=tile_mode)
    for key, sl in (("head", slice(0, 512)), ("transition", slice(46000, 50096)), ("strided", slice(None, None, 37))):
        err = np.max(np.abs(y[sl] - gold[key]))
        assert err <= TOL, f"tile_mode {tile_mode} {key}: {err:.3e}"


@pytest.mark.parametrize("name", ["a2_lite", "a2_full"])
@pytest.mark.parametrize("geom", [0, 2], ids=["auto", "tile512"])
def test_a2_family_lockstep(name, geom):
    """A2: look-backs of up to 14 x 239 frames span 7..14 tiles; the kernel-16 head convolution hands its
    accumulator columns over the same way."""
    nam = fx.load_model(name)
    x = fx.synthetic_batch(2, 9000, seed=5)
    ref = _oracle(nam, x)
    got = _gpu_chunks(nam, x, [4096, 300, 4604], 4700, kernel_geometry=geom)
    err = np.max(np.abs(got - ref))
    assert err <= TOL, f"{name} geometry {geom}: {err:.3e}"


def test_small_models_and_partial_last_tile():
    """wavenet.nam (3 / 2 channels padded to 4) and a single-array net; call lengths that leave 1 and 255 frames in
    the last tile."""
    for nam in (fx.load_model("wavenet"), fx.random_wavenet(channels=(8,), dilations=[[1, 2, 4, 8, 16, 32, 64, 128]], seed=3)):
        x = fx.synthetic_batch(2, 3000, seed=9)
        ref = _oracle(nam, x)
        for chunks in ([257, 511, 2232], [1025, 1975]):
            got = _gpu_chunks(nam, x, chunks, 2300)
            assert np.max(np.abs(got - ref)) <= TOL


def test_lockstep_equals_serial_walk_bitwise():
    """Same arithmetic in the same order: the launch mode must not change a single bit."""
    nam = fx.load_model("wavenet_a1_standard")
    x = fx.synthetic_batch(2, 8192, seed=2)
    a = _gpu_chunks(nam, x, [4096], 4096, fast_tanh=True, tile_mode=0)
    b = _gpu_chunks(nam, x, [4096], 4096, fast_tanh=True, tile_mode=2)
    assert np.array_equal(a, b)
