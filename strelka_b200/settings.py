"""SettingsManager mirror (reference: include/settings/settings.h:11-63).

String-keyed, string-valued map with typed setAs/getAs, same keys as the reference
(SURVEY.md section 5).  A missing key raises KeyError (the reference assert(0)s, settings.h:17-24).
"""
from __future__ import annotations

from ._abi import sb_settings


class SettingsManager:
    def __init__(self):
        self._map: dict[str, str] = {}

    def setAs(self, name: str, value) -> None:  # noqa: N802 (reference naming)
        if isinstance(value, bool):
            self._map[name] = "1" if value else "0"
        else:
            self._map[name] = str(value)

    def getAs(self, typ, name: str):  # noqa: N802
        if name not in self._map:
            raise KeyError(f"The setting {name} does not exist")
        v = self._map[name]
        if typ is bool:
            return v not in ("0", "", "False", "false")
        if typ is int:
            return int(float(v))
        return typ(v)

    def has(self, name: str) -> bool:
        return name in self._map

    # -- the keys OptiXRender::render reads (OptixRender.cpp:910-1004) -> POD for the C ABI
    def to_sb_settings(self) -> sb_settings:
        s = sb_settings()
        g = self.getAs
        s.spp = g(int, "render/pt/spp")
        s.spp_total = g(int, "render/pt/sppTotal")
        s.depth = g(int, "render/pt/depth")
        s.enable_acc = 1 if g(bool, "render/pt/enableAcc") else 0
        s.rect_light_sampling_method = g(int, "render/pt/rectLightSamplingMethod")
        s.debug = g(int, "render/pt/debug")
        s.shadow_ray_tmin = g(float, "render/pt/dev/shadowRayTmin")
        s.material_ray_tmin = g(float, "render/pt/dev/materialRayTmin")
        s.tonemapper_type = g(int, "render/pt/tonemapperType")
        s.gamma = g(float, "render/post/gamma")
        s.film_iso = g(float, "render/post/tonemapper/filmIso")
        s.cm2_factor = g(float, "render/post/tonemapper/cm2_factor")
        s.f_stop = g(float, "render/post/tonemapper/fStop")
        s.shutter_speed = g(float, "render/post/tonemapper/shutterSpeed")
        # extensions (no reference key): multi-GPU sample-stride sharding
        s.sample_offset = g(int, "render/b200/sampleOffset") if self.has("render/b200/sampleOffset") else 0
        s.sample_stride = g(int, "render/b200/sampleStride") if self.has("render/b200/sampleStride") else 1
        return s


def default_settings(spp_total: int = 64, spp: int = 1) -> SettingsManager:
    """The defaults the reference app installs (src/hdRunner/main.cpp:510-542)."""
    m = SettingsManager()
    m.setAs("render/width", 1024)
    m.setAs("render/height", 768)
    m.setAs("render/pt/depth", 4)
    m.setAs("render/pt/sppTotal", spp_total)
    m.setAs("render/pt/spp", spp)
    m.setAs("render/pt/tonemapperType", 0)
    m.setAs("render/pt/debug", 0)
    m.setAs("render/pt/enableAcc", True)
    m.setAs("render/pt/rectLightSamplingMethod", 0)
    m.setAs("render/post/tonemapper/filmIso", 100.0)
    m.setAs("render/post/tonemapper/cm2_factor", 1.0)
    m.setAs("render/post/tonemapper/fStop", 4.0)
    m.setAs("render/post/tonemapper/shutterSpeed", 100.0)
    m.setAs("render/post/gamma", 2.4)
    m.setAs("render/pt/dev/shadowRayTmin", 0.0)
    m.setAs("render/pt/dev/materialRayTmin", 0.0)
    return m
