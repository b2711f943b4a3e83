"""Seeded synthetic scene generators for the BASELINE.json configs.

Each generator fills a strelka_b200.Scene exactly the way the Hydra delegate would flatten the
corresponding USD stage (un-indexed triangle soup per mesh, packed normals/tangents, one mesh copy
per instance, curves with phantom points) and returns (scene, settings, (width, height)).
"""
from .cornell import make_cornell  # noqa: F401
from .hair import make_hair  # noqa: F401
from .kitchen import make_kitchen  # noqa: F401
from .instanced import make_instanced  # noqa: F401
from .common import make_quad_mesh, make_box_mesh, make_icosphere, tangent_for  # noqa: F401
