#!/usr/bin/env python3
"""BASELINE.json configs[3]: 3840 x 2160 RGB8 frames of a dense 24 x 13 board (device-resident),
full detect.  usage: python tools/config4_4k.py [n_frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = 3840, 2160
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
det.set_option("max_saddles", 4096)
det.set_option("chunk_frames", 16)
gray = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
det.render_boards_device(gray.data_ptr(), n, W, H, 24, 13, 4000)
rgb = gray[..., None].expand(n, H, W, 3).contiguous()
cap = 512
tags = torch.zeros((n, cap * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
for fmt, fr, name in ((pkg.FMT_RGB8, rgb, "RGB8"), (pkg.FMT_L8, gray, "L8")):
    for _ in range(2):
        det.detect_batch_device(fr.data_ptr(), n, W, H, fmt, tags.data_ptr(), cap, cnt.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        det.detect_batch_device(fr.data_ptr(), n, W, H, fmt, tags.data_ptr(), cap, cnt.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("4K %s dense board 24x13: %d frames in %.1f ms = %.0f frames/s; tags/frame %.1f, status bits %s"
          % (name, n, dt * 1e3, n / dt, float(cnt.float().mean()), sorted(set(st.cpu().tolist()))))
det.close()
