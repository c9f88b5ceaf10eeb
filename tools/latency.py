#!/usr/bin/env python3
"""Single-frame latency of TagDetector.detect (host image in, tags on host out) for several
warps-per-frame settings, and the 8-camera-rig shape (2048x1536).  usage: python tools/latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
for (w, h) in ((1280, 1024), (2048, 1536)):
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    d_img = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
    det.render_boards_device(d_img.data_ptr(), 1, w, h, 6, 6, 77)
    torch.cuda.synchronize()
    img = d_img[0].cpu().numpy()
    for warps in (0, 1, 2, 4, 8):
        det.set_option("board_warps", warps)
        for _ in range(3):
            tags = det.detect(img)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            tags = det.detect(img)
            ts.append(time.perf_counter() - t0)
        print("%dx%d board_warps=%d: detect latency median %.3f ms  min %.3f ms  (%d tags)"
              % (w, h, warps, 1e3 * np.median(ts), 1e3 * min(ts), len(tags)))
    det.close()
