#!/usr/bin/env python3
"""Per-frame latency distribution and phase breakdown of the board kernel (K6).
usage: python tools/board_times.py [n_frames] [board_warps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
warps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
W, H = 1280, 1024
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
det.set_option("chunk_frames", n)
det.set_option("board_warps", warps)
det.set_option("board_timing", 1)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 1000, stream=s.cuda_stream)
tags = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
st = torch.zeros(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, tags.data_ptr(), 64, cnt.data_ptr(),
                            st.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
t = det._board_times(0, n).astype(np.float64)
ns = t[:, 0]
print("frames %d warps/frame %d: board-kernel latency per frame (us): mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f"
      % (n, warps, ns.mean() / 1e3, np.percentile(ns, 50) / 1e3, np.percentile(ns, 90) / 1e3,
         np.percentile(ns, 99) / 1e3, ns.max() / 1e3))
names = {1: "grid+seeds", 2: "seed begin (50-NN, classify)", 3: "enumerate quads", 4: "score quads (groups)",
         5: "redo (general path)", 6: "rebuild winner + fix_missing"}
cyc = t[:, 1:7]
tot = cyc.sum(axis=1)
print("phase cycles per frame (mean, share of instrumented):")
for k, nm in names.items():
    print("  %-30s %10.0f  %5.1f%%" % (nm, t[:, k].mean(), 100.0 * t[:, k].sum() / tot.sum()))
print("  seeds processed %.1f  quads scored %.1f  redo batches %.2f  saddles %.0f  fast %.2f"
      % (t[:, 8].mean(), t[:, 9].mean(), t[:, 10].mean(), t[:, 11].mean(), (t[:, 12].astype(np.uint32) >> 31).mean()))
print("  decode + removal of used saddles: %.0f cycles per frame" % (t[:, 12].astype(np.uint32) & 0x7fffffff).mean())
print("  lockstep iterations %.0f  expansion attempts %.0f (%.2f groups / iteration)  tuple passes %.0f"
      % (t[:, 13].mean(), t[:, 14].mean(), t[:, 14].sum() / max(t[:, 13].sum(), 1), t[:, 15].mean()))
print("  lockstep cycles: take-work %.0f  advance %.0f  queries %.0f  tuples %.0f  (per frame, warp 0)"
      % (t[:, 16].mean(), t[:, 17].mean(), t[:, 18].mean(), t[:, 19].mean()))
print("  searches: lanes asking %.0f  lanes searching (cache miss) %.0f = %.1f%%  iterations with a search %.0f"
      % (t[:, 20].mean(), t[:, 21].mean(), 100.0 * t[:, 21].sum() / max(t[:, 20].sum(), 1), t[:, 22].mean()))
if t[:, 24].sum() > 0:  # only in builds with -DAGB_LOOP_STATS
    print("  search loops: bucket rows %.0f  candidate iterations %.0f  (%.1f lanes busy per iteration)"
          % (t[:, 23].mean(), t[:, 24].mean(), t[:, 25].sum() / max(t[:, 24].sum(), 1)))
worst = np.argsort(-ns)[:5]
for f in worst[:3]:
    print("  slow frame %d: %.0f us, seeds %d quads %d redo %d saddles %d tags %d; cycles %s"
          % (f, ns[f] / 1e3, t[f, 8], t[f, 9], t[f, 10], t[f, 11], int(cnt[f]), t[f, 1:7].astype(int).tolist()))
det.close()
