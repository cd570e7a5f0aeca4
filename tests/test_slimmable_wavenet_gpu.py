"""GPU parity of "slimmable" WaveNets (NAM/wavenet/model.cpp:1290-1315 -> NAM/wavenet/slimmable.cpp):
SetSlimmableSize(ratio) on the handle must behave like the reference's SlimmableWavenet -- a freshly reset, prewarmed
WaveNet built from the leading channels of every tensor.  The sub-model documents are pinned against the reference
build on CPU (tests/test_reference_build.py::test_slimmable_wavenet_slicing); here the CUDA path is held to the oracle
running those documents.  Same 1e-5 gate (relative to max(1, |y|): these random-weight nets reach |y| ~ 10)."""
import numpy as np
import pytest

import neuralampmodelercore_b200 as nb
from oracle import oracle
from tests import nam_fixtures as fx
from tests.test_reference_build import _slimmable_models

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("case", [0, 1])
def test_set_slimmable_size(case):
    label, nam = list(_slimmable_models())[case]
    subs = nb.submodels(nam)
    d = nb.get_dsp(nam, batch=3)
    assert d.GetSlimmableSizeBreakpoints() == [mv for mv, _ in subs[:-1]]  # get_ratio_breakpoints, slimmable.cpp:111-125
    d.Reset(48000.0, 256)
    x = fx.synthetic_batch(3, 1024, seed=8)
    bps = [mv for mv, _ in subs[:-1]]
    for val in [1.0, 0.0] + [(a + b) / 2 for a, b in zip([0.0] + bps, bps + [1.0])] + bps:
        d.SetSlimmableSize(val)  # the newly active sub-model starts reset + prewarmed (slimmable.cpp:455-475)
        got = np.concatenate([d.process_batch(np.ascontiguousarray(x[:, p:p + 256])) for p in range(0, 1024, 256)], axis=1)
        idx = next((i for i, (mv, _) in enumerate(subs) if val < mv), len(subs) - 1)
        o = oracle.OracleModel.from_dict(subs[idx][1])
        o.reset(48000.0, 64)
        ref = o.run_batch(x, 64)
        o.close()
        err = float(np.max(np.abs(got - ref))) / max(1.0, float(np.max(np.abs(ref))))
        assert err <= TOL, f"{label}: ratio {val} (sub-model {idx}): {err:.3e}"
        d.SetSlimmableSize(1.0 if idx != len(subs) - 1 else 0.0)  # leave, so the next visit starts fresh again
    d.close()


def test_default_is_full_size_and_plain_wavenets_are_not_slimmable():
    nam = fx.load_model("slimmable_wavenet")
    x = fx.synthetic_batch(1, 640, seed=3)
    d = nb.get_dsp(nam)
    d.Reset(48000.0, 64)
    got = np.concatenate([d.process_batch(x[:, p:p + 64]) for p in range(0, 640, 64)], axis=1)
    d.close()
    full = dict(nam)
    full["config"] = {**nam["config"], "layers": [{**lc, "slimmable": None} for lc in nam["config"]["layers"]]}
    o = oracle.OracleModel.from_dict(full)
    o.reset(48000.0, 64)
    ref = o.run(x[0], 64)
    assert float(np.max(np.abs(got[0] - ref))) / max(1.0, float(np.max(np.abs(ref)))) <= TOL
    plain = nb.get_dsp(full)
    assert plain.GetSlimmableSizeBreakpoints() == []
    with pytest.raises(nb.UnsupportedModelError):
        plain.SetSlimmableSize(0.5)
    plain.close()
