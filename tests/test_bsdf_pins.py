"""Independent pins of the CPU oracle's material models: values against numpy restatements of the
published formulas (tests/bsdf_ref.py shares no code with oracle/bsdf.h, oracle/hair.h or the device files), pdf
normalisation, sample/evaluate consistency, sampled histogram vs pdf, energy (one-bounce white furnace), reciprocity.
The device implementations get the same checks in tests/test_gpu_bsdf_pins.py."""
import numpy as np
import pytest

import bsdf_pins_common as P
from oracle import pyoracle
from strelka_b200 import _abi


def run(m, packed):
    out = np.zeros((len(packed), 15), dtype=np.float32)
    pyoracle.lib().orc_bsdf_batch(m.ctypes.data, len(packed), packed.ctypes.data, out.ctypes.data)
    return out[:, :8], out[:, 8:]


ALL = ([(_abi.SB_MATERIAL_DIFFUSE, dict(base_color=(0.7, 0.4, 0.2)))] + [(_abi.SB_MATERIAL_USD_PREVIEW_SURFACE, kw) for kw in P.UPS_CASES]
       + [(_abi.SB_MATERIAL_HAIR, kw) for kw in P.HAIR_CASES])
IDS = [f"{m}-{i}" for i, (m, _) in enumerate(ALL)]


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_evaluate_equals_published_formulas(model, kw):
    P.check_values_against_reference(run, model, kw)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_sample_and_evaluate_agree(model, kw):
    P.check_sample_evaluate_consistency(run, model, kw)


@pytest.mark.parametrize("model,kw", ALL, ids=IDS)
def test_pdf_integrates_to_one_and_weights_carry_the_albedo(model, kw):
    k1 = (0.3, -0.2, 0.8) if model != _abi.SB_MATERIAL_HAIR else (0.35, 0.7, -0.4)
    P.check_pdf_normalisation_and_energy(run, model, kw, k1)


@pytest.mark.parametrize("model,kw", [ALL[0], ALL[3], ALL[7], ALL[9]], ids=["lambert", "ups-r0.5", "ups-clearcoat", "hair"])
def test_sampled_directions_follow_the_pdf(model, kw):
    k1 = (0.5, 0.1, 0.6) if model != _abi.SB_MATERIAL_HAIR else (-0.3, 0.8, 0.2)
    P.check_sampling_matches_pdf(run, model, kw, k1)


def test_white_furnace_one_bounce():
    # white Lambert and absorption-free hair return all the light they receive; a white metal (F0 = 1) loses only what
    # single-scattering GGX loses, never gains
    P.check_pdf_normalisation_and_energy(run, _abi.SB_MATERIAL_DIFFUSE, dict(base_color=(1, 1, 1)), (0.1, 0.4, 0.7), white=True)
    for k1 in ((0.2, 0.9, 0.1), (0.8, 0.3, -0.3), (0.0, -0.6, 0.8)):
        P.check_pdf_normalisation_and_energy(run, _abi.SB_MATERIAL_HAIR, P.WHITE_HAIR, k1, white=True)
    k1 = P.unit((0.4, 0.0, 0.9))
    prev = 1.01
    for rough in (0.1, 0.5, 1.0):
        a = P.check_pdf_normalisation_and_energy(run, _abi.SB_MATERIAL_USD_PREVIEW_SURFACE,
                                                 dict(base_color=(1, 1, 1), roughness=rough, metallic=1.0), k1)
        assert np.all(a < prev)  # rougher single-scattering GGX loses more
        prev = a.max()
    # closed form at alpha = 1 (D = 1/pi, G1 = 2c/(1+c), F = 1): albedo = 2 (1 - ln 2) / (1 + cos theta_1)
    assert abs(prev - 2.0 * (1.0 - np.log(2.0)) / (1.0 + k1[2])) < 1e-3


def test_reciprocity_of_the_single_lobes():
    # the layered UsdPreviewSurface is not reciprocal by construction (weights depend on the view cosine only, like
    # MDL's fresnel_layer); its building blocks are
    P.check_reciprocity(run, _abi.SB_MATERIAL_DIFFUSE, dict(base_color=(0.3, 0.6, 0.9)))
    P.check_reciprocity(run, _abi.SB_MATERIAL_USD_PREVIEW_SURFACE, dict(base_color=(0.9, 0.6, 0.3), roughness=0.5, metallic=1.0))
