"""The independent material pins of tests/test_bsdf_pins.py, run against the DEVICE implementations through
sb_test_bsdf (the functions the shade kernel calls): values vs numpy restatements of the published formulas, pdf
normalisation, sample/evaluate consistency, histogram vs pdf, one-bounce white furnace.  Plus device == oracle."""
import numpy as np
import pytest

import bsdf_pins_common as P
from strelka_b200 import _abi
from test_bsdf_pins import ALL, IDS
from test_bsdf_pins import run as run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run(gpu_render):
    return lambda m, packed: gpu_render.test_bsdf(m, packed)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_device_evaluate_equals_published_formulas(run, model, kw):
    P.check_values_against_reference(run, model, kw)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_device_sample_and_evaluate_agree(run, model, kw):
    P.check_sample_evaluate_consistency(run, model, kw)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_device_pdf_integrates_to_one_and_weights_carry_the_albedo(run, model, kw):
    k1 = (0.3, -0.2, 0.8) if model != _abi.SB_MATERIAL_HAIR else (0.35, 0.7, -0.4)
    P.check_pdf_normalisation_and_energy(run, model, kw, k1)


@pytest.mark.parametrize("model,kw", [ALL[0], ALL[3], ALL[7], ALL[9]], ids=["lambert", "ups-r0.5", "ups-clearcoat", "hair"])
def test_device_sampled_directions_follow_the_pdf(run, model, kw):
    k1 = (0.5, 0.1, 0.6) if model != _abi.SB_MATERIAL_HAIR else (-0.3, 0.8, 0.2)
    P.check_sampling_matches_pdf(run, model, kw, k1)


def test_device_white_furnace_one_bounce(run):
    P.check_pdf_normalisation_and_energy(run, _abi.SB_MATERIAL_DIFFUSE, dict(base_color=(1, 1, 1)), (0.1, 0.4, 0.7), white=True)
    for k1 in ((0.2, 0.9, 0.1), (0.8, 0.3, -0.3)):
        P.check_pdf_normalisation_and_energy(run, _abi.SB_MATERIAL_HAIR, P.WHITE_HAIR, k1, white=True)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_device_matches_oracle(run, model, kw):
    """same inputs through both implementations (hair: float kernels vs the oracle's double-precision restatement)"""
    rng = np.random.default_rng(11)
    n = 20000
    hair = model == _abi.SB_MATERIAL_HAIR
    k1 = P.unit(rng.normal(size=(n, 3)))
    if not hair:
        k1[:, 2] = np.abs(k1[:, 2]) * 0.98 + 0.02
        k1 = P.unit(k1)
    fr = (P.N_HAIR, P.N_HAIR, P.T_HAIR) if hair else (P.N_SURF, P.N_SURF, P.T_SURF)
    packed = P.pack(*fr, k1, rng.random((n, 4)), P.unit(rng.normal(size=(n, 3))))
    m = P.material(model, **kw)
    sg, eg = run(m, packed)
    so, eo = run_oracle(m, packed)
    same_event = sg[:, 7] == so[:, 7]
    assert same_event.mean() > 0.9995  # a lobe-selection threshold may fall between the float and the double value
    tol = dict(rtol=2e-3, atol=2e-4) if hair else dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(sg[same_event, :7], so[same_event, :7], **tol)
    np.testing.assert_allclose(eg, eo, **(dict(rtol=1e-3, atol=1e-5) if hair else dict(rtol=2e-4, atol=1e-6)))
