#!/usr/bin/env python3
"""e2e (pinned host frames in, tags out) for several host chunk sizes / warps per frame."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package()
B, W, H, cap = 1024, 1280, 1024, 64
det0 = pkg.TagDetector(pkg.TagFamily.T36H11)
frames = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
det0.render_boards_device(frames.data_ptr(), B, W, H, 6, 6, 1000)
torch.cuda.synchronize()
det0.close()
h = torch.empty((B, H, W), dtype=torch.uint8).pin_memory(); h.copy_(frames); torch.cuda.synchronize()
hf = h.numpy()
out = torch.zeros((B, cap * 9), dtype=torch.int32).pin_memory().numpy().view(pkg.TAG_DTYPE).reshape(B, cap)
cnt = torch.zeros(B, dtype=torch.int32).pin_memory().numpy()
st = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
for hc, bw in [(128, 0), (128, 2), (128, 4), (64, 2), (64, 4), (256, 2), (96, 2)]:
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    det.set_option("host_chunk_frames", hc)
    det.set_option("board_warps", bw)
    for _ in range(2): det.detect_batch_into(hf, out, cnt, st)
    t0 = time.perf_counter()
    for _ in range(5): det.detect_batch_into(hf, out, cnt, st)
    dt = time.perf_counter() - t0
    print("host chunk %3d warps %d: e2e %.0f frames/s (%.1f GB/s H2D)" % (hc, bw, 5 * B / dt, 5 * B * W * H / dt / 1e9))
    det.close()
