#!/usr/bin/env python3
"""K6 (board search + decode) timed alone under different settings: how its launch time depends on
the boards searched per frame, the resident blocks per SM (board_smem_pad) and the launch size.
usage: python tools/k6_probe.py [n_frames] ; env AG_LIB selects another build of the library."""
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
torch.cuda.set_stream(s)
frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
tags = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")


def run(label, max_boards=2, pad=0, warps=0, chunk=None, reps=3, opts=()):
    det = pkg.TagDetector(pkg.TagFamily.T36H11, pkg.DetectorParams(max_num_of_boards=max_boards))
    det.set_option("chunk_frames", chunk or n)
    det.set_option("board_warps", warps)
    if pad:
        det.set_option("board_smem_pad", pad)
    for k, v in opts:
        det.set_option(k, v)
    det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 1000, stream=s.cuda_stream)
    for r in range(reps + 1):
        if r == 1:
            torch.cuda.synchronize()
            det.stage_times(reset=True)
            det.set_option("profile", 1)
        det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, tags.data_ptr(), 64, cnt.data_ptr(),
                                st.data_ptr(), stream=s.cuda_stream)
        torch.cuda.synchronize()
    t = det.stage_times(reset=True)
    per = {k: 1024.0 * v[0] / (reps * n) for k, v in t.items()}
    print("%-34s K6 %.2f  K1 %.2f K2 %.2f K3 %.2f K4 %.2f ms/1024 frames; tags/frame %.2f"
          % (label, per["boards_decode"], per["blur_hessian_min"], per["threshold"], per["label_centroid"],
             per["refine_filter"], float(cnt.float().mean())), flush=True)
    if pad:
        det.set_option("board_smem_pad", 0)
    det.close()


which = os.environ.get("K6_PROBE", "all")
run("default (2 boards, n=%d)" % n)
if which == "rounds":
    run("1 board", max_boards=1)
if which == "split":
    run("board_split = 0", opts=(("board_split", 0),))
    run("4 warps / frame", warps=4)
    run("1 board", max_boards=1)
if which == "all":
    run("1 board", max_boards=1)
    # 37.1 KB per block -> 6 blocks per SM; pads chosen so that 5, 4, 3, 2 blocks fit in 227 KB
    if os.environ.get("AG_LIB"):  # board_smem_pad exists in experiment builds only
        for blocks, pad in ((5, 8 * 1024), (4, 19 * 1024), (3, 38 * 1024), (2, 76 * 1024)):
            run("%d blocks / SM" % blocks, pad=pad)
    run("4 warps / frame", warps=4)
    for c in (296, 444, 888, 1024):
        run("chunk %d (sync calls)" % c, chunk=c)
