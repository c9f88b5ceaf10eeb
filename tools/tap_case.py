#!/usr/bin/env python3
"""One fixture image through the stage-tap path (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from PIL import Image
import __graft_entry__ as entry
pkg = entry.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "EuRoC"
img = np.ascontiguousarray(np.array(Image.open(os.path.join(ROOT, "tests/golden/images", name + ".png"))))
det = pkg.TagDetector(pkg.TagFamily.T36H11)
g = det.stages(img)
print(name, "tags", len(g["tags"]), "refined", len(g["refined"]), "quads", len(g["quads"]))
det.close()
