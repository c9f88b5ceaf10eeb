#!/usr/bin/env python3
"""detect_batch through ag_multi_* (one process, one host thread per GPU, frames in pinned host memory)
on every GPU of the box: frames/s end to end.  usage: python tools/multi_e2e.py [frames_per_gpu] [steps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W, H, cap = 1280, 1024, 64
pkg = entry.load_package()
G = torch.cuda.device_count()
n = per_gpu * G
single = pkg.TagDetector(pkg.TagFamily.T36H11, None, device=0)
d = torch.empty((per_gpu, H, W), dtype=torch.uint8, device="cuda:0")
single.render_boards_device(d.data_ptr(), per_gpu, W, H, 6, 6, 1000)
torch.cuda.synchronize()
frames = pkg.pinned_empty((n, H, W))
one = d.cpu().numpy()
for g in range(G):
    frames[g * per_gpu:(g + 1) * per_gpu] = one
out = np.zeros((n, cap), pkg.TAG_DTYPE)
cnt = np.zeros(n, np.int32)
st = np.zeros(n, np.uint32)
ref = (np.zeros((per_gpu, cap), pkg.TAG_DTYPE), np.zeros(per_gpu, np.int32), np.zeros(per_gpu, np.uint32))
single.detect_batch_into(one, *ref)
single.close()
multi = pkg.MultiTagDetector(pkg.TagFamily.T36H11, None)
multi.detect_batch_into(frames, out, cnt, st)
t0 = time.perf_counter()
for _ in range(steps):
    multi.detect_batch_into(frames, out, cnt, st)
dt = time.perf_counter() - t0
same = all(np.array_equal(cnt[g * per_gpu:(g + 1) * per_gpu], ref[1]) and
           np.array_equal(out[g * per_gpu:(g + 1) * per_gpu], ref[0]) for g in range(G))
print(json.dumps({"api": "ag_multi_detect_batch (synchronous calls)", "gpus": G, "frames_per_call": n, "steps": steps,
                  "frames_per_s": n * steps / dt, "identical_to_single_gpu": bool(same)}))
multi.close()
pkg.pinned_free(frames)
