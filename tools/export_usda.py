"""Write the `.usda` twin of a BASELINE config ("flattened arrays + .usda twin"), to be rendered by a
real Strelka build for cross-validation:  python tools/export_usda.py c2 cornell.usda
The big configs are written at full size (c3 ~ 2 M triangles -> a few hundred MB of text); pass --small to shrink."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strelka_b200 import usd_twin  # noqa: E402
from strelka_b200.scenes import make_cornell, make_hair, make_instanced, make_kitchen  # noqa: E402

if __name__ == "__main__":
    cfg, path = sys.argv[1], sys.argv[2]
    small = "--small" in sys.argv
    make = {
        "c2": lambda: make_cornell(1024, 1024, 256),
        "c3": lambda: make_kitchen(1920, 1080, 2048, **({"n_props": 40, "subdiv": 2} if small else {})),
        "c4": lambda: make_hair(1024, 1024, 1024, depth=6, **({"n_strands": 2000} if small else {})),
        "c5": lambda: make_instanced(3840, 2160, 4096, **({"n_instances": 64} if small else {})),
    }[cfg]
    scene, settings, (w, h) = make()
    usd_twin.write_usda(scene, path, settings, w, h)
    print(f"{path}: {os.path.getsize(path) / 1e6:.1f} MB, {len(scene.instances)} instances, {len(scene.lights)} lights")
