"""Whole-model pin of the oracle against THE REFERENCE ITSELF: oracle/_ref/libnam_ref.so is the unmodified
NeuralAmpModelerCore source tree compiled against a stand-in for its missing Eigen submodule
(oracle/eigen_shim, recipe `make -C oracle ref`).  The reference's own tests hold no whole-model vectors
(SURVEY.md 8c), so this is what turns the oracle's model-level reading of the reference -- weight order, prewarm,
head_scale, condition_dsp, containers, the A2 fast path -- from "restated" into "checked".

Runs on CPU.  Skipped where the library was not built (no reference tree at build time)."""
import numpy as np
import pytest

from oracle import oracle, ref
from tests import nam_fixtures as fx

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnam_ref.so not built (no /root/reference)")

# two independent fp32 implementations (different summation order inside the matrix products): the reference's
# authors accept 5e-5 between two of their own (tools/test/test_a2_fast.cpp:296-298); measured here <= 3e-6
TOL = 1e-5

MODELS = ["wavenet", "wavenet_a1_standard", "lstm", "wavenet_condition_dsp", "wavenet_a2_max", "a2_lite", "a2_full"]


def _rel(err, y):
    return err / max(1.0, float(np.max(np.abs(y))))


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("fast", [False, True], ids=["exact_tanh", "fast_tanh"])
def test_oracle_matches_the_reference_build(name, fast):
    nam = fx.load_model(name)
    x = fx.input_wav()[43000:53000]  # 5,000 frames of silence, then the 220 Hz sine
    r = ref.ReferenceModel.from_dict(nam, fast_tanh=fast)
    r.reset(48000.0, 64)
    m = oracle.OracleModel.from_dict(nam, fast_tanh=fast)
    m.reset(48000.0, 64)
    assert r.prewarm_samples == m.prewarm_samples
    yr, yo = r.run(x, 64), m.run(x, 64)
    err = float(np.max(np.abs(yr - yo)))
    assert _rel(err, yr) <= TOL, f"{name}: max|reference - oracle| = {err:.3e} (max|y| {np.max(np.abs(yr)):.3f})"
    r.close()


@pytest.mark.parametrize("name", ["wavenet_a1_standard", "a2_full", "lstm"])
def test_block_size_protocols_agree(name):
    """The reference at maxBufferSize 64 (tools) vs 1000: same stream (prewarmed WaveNet state is block-size
    independent; the LSTM prewarm rounds up to whole blocks, lstm.cpp:127-134, so compare like with like)."""
    nam = fx.load_model(name)
    x = fx.synthetic_batch(1, 3000, seed=17)[0]
    for block in (64, 1000):
        r = ref.ReferenceModel.from_dict(nam)
        r.reset(48000.0, block)
        m = oracle.OracleModel.from_dict(nam)
        m.reset(48000.0, block)
        yr, yo = r.run(x, block), m.run(x, block)
        assert _rel(float(np.max(np.abs(yr - yo))), yr) <= TOL, f"{name} block {block}"
        r.close()


def test_golden_vectors_were_not_a_misreading():
    """The committed oracle vectors (tests/golden/oracle_outputs.npz) against the reference build, full 2 s."""
    x = fx.input_wav()
    for name in ("wavenet", "wavenet_a1_standard", "a2_full"):
        gold = fx.oracle_golden(name, "exact")
        r = ref.ReferenceModel.from_dict(fx.load_model(name))
        r.reset(48000.0, 64)
        y = r.run(x, 64)
        r.close()
        for key, val in fx.decimate(y).items():
            assert np.max(np.abs(val - gold[key])) <= TOL, f"{name} {key}"


def test_a2_fast_path_and_generic_path_of_the_reference():
    """The default reference build routes A2 shapes to A2FastModel (model.cpp:1317-1320); the second build has the
    fast path compiled out.  Both must agree with the oracle (and so with each other)."""
    if not ref.available("generic"):
        pytest.skip("generic variant not built")
    x = fx.synthetic_batch(1, 4000, seed=3)[0]
    for name in ("a2_lite", "a2_full"):
        nam = fx.load_model(name)
        m = oracle.OracleModel.from_dict(nam)
        m.reset(48000.0, 64)
        yo = m.run(x, 64)
        for variant in ("default", "generic"):
            r = ref.ReferenceModel.from_dict(nam, variant=variant)
            r.reset(48000.0, 64)
            err = float(np.max(np.abs(r.run(x, 64) - yo)))
            assert err <= TOL, f"{name} {variant}: {err:.3e}"
            r.close()


def test_slimmable_container_semantics():
    """ContainerModel (container.cpp): full size by default; SetSlimmableSize before or after Reset; the newly
    active sub-model starts from its own prewarmed state."""
    lite, full = fx.load_model("a2_lite"), fx.load_model("a2_full")
    cont = fx.make_container([(0.5, lite), (1.0, full)])
    x = fx.synthetic_batch(1, 2000, seed=9)[0]

    def orc(nam, sig):
        m = oracle.OracleModel.from_dict(nam)
        m.reset(48000.0, 64)
        return m.run(sig, 64)

    r = ref.ReferenceModel.from_dict(cont)
    r.reset(48000.0, 64)
    assert np.max(np.abs(r.run(x[:1000], 64) - orc(full, x[:1000]))) <= TOL
    r.set_slimmable_size(0.2)
    assert np.max(np.abs(r.run(x, 64) - orc(lite, x))) <= TOL
    r.set_slimmable_size(0.5)  # 0.5 is not < 0.5 -> full again, freshly reset
    assert np.max(np.abs(r.run(x[:640], 64) - orc(full, x[:640]))) <= TOL
    r.close()
    r = ref.ReferenceModel.from_dict(cont)
    r.set_slimmable_size(0.1)  # before the first Reset (tools/render.cpp:117-126 order)
    r.reset(48000.0, 64)
    assert np.max(np.abs(r.run(x, 64) - orc(lite, x))) <= TOL
    r.close()
