#!/usr/bin/env python3
"""Stage times of BASELINE.json configs[3] (3840 x 2160 RGB8, 24 x 13 board): CUDA events per stage.
usage: python tools/config4_stages.py [n_frames] [chunk]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W, H = 3840, 2160
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
if chunk:
    det.set_option("chunk_frames", chunk)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
gray = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
det.render_boards_device(gray.data_ptr(), n, W, H, 24, 13, 4000, stream=s.cuda_stream)
rgb = gray[..., None].expand(n, H, W, 3).contiguous()
del gray
cap = 512
tags = torch.zeros((n, cap * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
for r in range(3):
    if r == 1:
        torch.cuda.synchronize()
        det.stage_times(reset=True)
        det.set_option("profile", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    det.detect_batch_device(rgb.data_ptr(), n, W, H, pkg.FMT_RGB8, tags.data_ptr(), cap, cnt.data_ptr(), st.data_ptr(),
                            stream=s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
t = det.stage_times(reset=True)
print("%d frames: %.1f ms per call = %.0f frames/s; tags/frame %.1f" % (n, ms, n / ms * 1e3, float(cnt.float().mean())))
print({k: round(v[0] / 2, 2) for k, v in t.items()}, "ms per call (sum of launches)")
det.close()
