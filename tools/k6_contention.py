#!/usr/bin/env python3
"""How much does the board kernel (K6) slow down when other work shares the GPU?  Times K6 of one
2048-frame detect call alone, beside a pure HBM copy stream, and beside an L2-resident arithmetic
stream (torch kernels on a second stream).  usage: python tools/k6_contention.py [n_frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
W, H = 1280, 1024
pkg = entry.load_package()
s = torch.cuda.Stream()
side = torch.cuda.Stream()
torch.cuda.set_stream(s)
frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
tags = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
det = pkg.TagDetector(pkg.TagFamily.T36H11)
det.set_option("chunk_frames", n)
det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 1000, stream=s.cuda_stream)
big_a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
big_b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
small = torch.rand(1 << 22, device="cuda")  # 16 MB: stays in L2


def hog(kind, reps):
    with torch.cuda.stream(side):
        for _ in range(reps):
            if kind == "hbm":
                big_b.copy_(big_a)
            elif kind == "alu":
                small.mul_(1.0000001).add_(1e-9).mul_(0.9999999).sub_(1e-9)


def run(label, kind=None):
    for r in range(3):
        torch.cuda.synchronize()
        if r == 1:
            det.stage_times(reset=True)
            det.set_option("profile", 1)
        if kind:
            hog(kind, 400 if kind == "hbm" else 4000)  # well over the call's duration
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, tags.data_ptr(), 64, cnt.data_ptr(),
                                st.data_ptr(), stream=s.cuda_stream)
        e1.record(s)
        e1.synchronize()
        if r == 2:
            ms = e0.elapsed_time(e1)
        torch.cuda.synchronize()
    det.set_option("profile", 0)
    t = det.stage_times(reset=True)
    per = {k: 1024.0 * v[0] / (2 * n) for k, v in t.items()}
    print("%-28s call %.2f ms | per 1024 frames: K6 %.2f  K1 %.2f K2 %.2f K3 %.2f K4 %.2f"
          % (label, ms, per["boards_decode"], per["blur_hessian_min"], per["threshold"], per["label_centroid"],
             per["refine_filter"]), flush=True)


run("alone")
run("beside an HBM copy stream", "hbm")
run("beside an L2/ALU stream", "alu")
run("alone again")
det.close()
