#!/usr/bin/env python3
"""Build the committed golden fixtures (run in the build container, where /root/reference exists).

* copies the reference's own test images (tests/data/*.png, data/*.png: the inputs its
  integration tests and benches use, tests/test_detector.rs:26-32) to tests/golden/images/
  so that the GPU box, which has no /root/reference, can run the same inputs;
* records the expected tag COUNTS the reference asserts (tests/test_detector.rs:26-32) and,
  next to them, what the CPU oracle returns for every image (ids, corners, stage statistics),
  in tests/golden/expected.json.  The oracle values are a regression pin for the oracle
  itself; the counts are the reference's pins.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"
REFERENCE_COUNTS = {  # tests/test_detector.rs:26-32
    "iphone": 66, "EuRoC": 36, "TUM_VI": 36, "right": 36, "r45": 36, "top": 36, "two_boards": 72,
}
IMAGES = {n: "tests/data/%s.png" % n for n in list(REFERENCE_COUNTS) + ["top_right"]}
IMAGES["demo_1520525725372653511"] = "data/1520525725372653511.png"


def main():
    os.makedirs(os.path.join(HERE, "images"), exist_ok=True)
    out = {}
    for name, rel in IMAGES.items():
        dst = os.path.join(HERE, "images", name + ".png")
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
        img = np.array(Image.open(dst))
        fe = oracle.front_end(img, want_labels=True)
        tags = oracle.detect(img)
        out[name] = {
            "source": rel,
            "shape": list(img.shape), "dtype": str(img.dtype),
            "reference_count": REFERENCE_COUNTS.get(name),
            "oracle": {
                "count": len(tags),
                "ids": sorted(tags),
                "corners": {str(k): [[float(v) for v in p] for p in tags[k]] for k in sorted(tags)},
                "min_response_bits": int(np.float32(fe["min"]).view(np.uint32)),
                "n_clusters": int(len(fe["centers"])),
                "n_raw": int(len(fe["raw"])),
                "n_refined": int(len(fe["refined"])),
                "mask_pixels": int((fe["labels"] >= 0).sum()),
                "blur_sha256": hashlib.sha256(fe["blur"].tobytes()).hexdigest(),
                "resp_sha256": hashlib.sha256(fe["resp"].tobytes()).hexdigest(),
                "labels_sha256": hashlib.sha256(fe["labels"].tobytes()).hexdigest(),
            },
        }
        print(name, img.shape, img.dtype, "ref", REFERENCE_COUNTS.get(name), "oracle", len(tags))
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
