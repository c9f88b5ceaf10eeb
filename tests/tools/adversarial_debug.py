#!/usr/bin/env python3
"""Which path disagrees with the oracle on the gate-adversarial saddle sets of
tests/test_gpu_parity.py (throughput path / general path of the board kernel)?  GPU needed."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402,F401
import oracle  # noqa: E402
import synth  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

spec = importlib.util.spec_from_file_location("tgp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tgp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tgp)
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
img = synth.render_board_numpy(640, 480, seed=3, tag_px=44.0)
base = oracle.front_end(img, want_labels=False)["refined"]
rng = np.random.default_rng(2024)
for name, s in synth.adversarial_saddle_sets(base, rng):
    want = oracle.try_find_best_board(s)
    want = np.zeros((0, 4), np.int32) if want is None else want
    res = {}
    for fast in (1, 0):
        det.set_option("board_fast", fast)
        res[fast], _ = det._boards_from_saddles(s, img)
    det.set_option("board_fast", 1)
    flags = {k: np.array_equal(v, want) for k, v in res.items()}
    if not all(flags.values()):
        print(name, "fast ok" if flags[1] else "FAST DIFFERS", "general ok" if flags[0] else "GENERAL DIFFERS",
              "n oracle %d fast %d general %d" % (len(want), len(res[1]), len(res[0])))
        ws = set(map(tuple, want.tolist()))
        for k, v in res.items():
            gs = set(map(tuple, v.tolist()))
            print("   fast=%d: only oracle %s | only gpu %s" % (k, sorted(ws - gs)[:4], sorted(gs - ws)[:4]))
        np.save(os.path.join(ROOT, "gpurun_out", "adv_%s.npy" % name), s)
print("done")
