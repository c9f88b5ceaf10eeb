#!/usr/bin/env python3
"""Run the device-batch detect path repeatedly on the same frames (streaming calls, all pipeline
slots in use) and check that every repetition gives bit-identical results, and that they equal the
CPU oracle on a sample.  usage: python tests/tools/stress_determinism.py [n_frames] [reps] [warps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
warps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, H = 1280, 1024
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
det.set_option("chunk_frames", 256)
det.set_option("board_warps", warps)
det.set_option("device_async", 1)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 4242, stream=s.cuda_stream)
outs = []
for r in range(reps):
    tags = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, tags.data_ptr(), 64, cnt.data_ptr(),
                            st.data_ptr(), stream=s.cuda_stream)
    outs.append((tags, cnt, st))
det.detect_batch_device_wait(stream=s.cuda_stream)
torch.cuda.synchronize()
bad = 0
t0, c0, s0 = outs[0]
for r in range(1, reps):
    t, c, st = outs[r]
    if not torch.equal(c, c0):
        d = (c != c0).nonzero().flatten().tolist()
        print("rep %d: counts differ at frames %s" % (r, d[:10]))
        bad += 1
    elif not torch.equal(t, t0):
        d = (t != t0).any(dim=1).nonzero().flatten().tolist()
        print("rep %d: tags differ at frames %s" % (r, d[:10]))
        bad += 1
print("repetitions %d, frames %d: %s; mean tags/frame %.3f, status bits %s"
      % (reps, n, "all identical" if bad == 0 else "%d MISMATCHING" % bad, float(c0.float().mean()),
         sorted(set(s0.cpu().numpy().tolist()))))
# oracle on a sample
oracle = entry.load_oracle()
k = int(os.environ.get("AG_ORACLE_SAMPLE", "16"))
sample = frames[:k].cpu().numpy()
want = oracle.detect_batch(sample)
tg = t0[:k].cpu().numpy().view(pkg.TAG_DTYPE).reshape(k, 64)
cn = c0[:k].cpu().numpy()
okc = 0
for i in range(k):
    got = {int(x["id"]): x["xy"].reshape(4, 2) for x in tg[i][:cn[i]]}
    if sorted(got) == sorted(want[i]) and all(np.abs(got[j] - want[i][j]).max() <= 1e-3 for j in got):
        okc += 1
    else:
        print("frame %d differs from the oracle: got %s want %s" % (i, sorted(got), sorted(want[i])))
print("oracle check: %d / %d frames identical" % (okc, k))
det.close()
sys.exit(1 if bad or okc != k else 0)
