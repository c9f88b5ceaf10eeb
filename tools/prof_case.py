#!/usr/bin/env python3
"""Small fixed workload for ncu: render N board frames on the device and run detect once.
usage: python tools/prof_case.py [n_frames] [chunk] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else n
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
W, H = 1280, 1024
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
det.set_option("chunk_frames", chunk)
if os.environ.get("AG_BOARD_GRID") is not None:
    det.set_option("board_grid", int(os.environ["AG_BOARD_GRID"]))
for key in ("board_warps", "board_fast", "board_lattice"):
    if os.environ.get("AG_" + key.upper()) is not None:
        det.set_option(key, int(os.environ["AG_" + key.upper()]))
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 1000, stream=s.cuda_stream)
tags = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
for _ in range(reps):
    det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, tags.data_ptr(), 64, cnt.data_ptr(),
                            st.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
print("frames", n, "tags/frame", float(cnt.float().mean()), "launches", det.launch_count)
det.close()
