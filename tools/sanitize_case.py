#!/usr/bin/env python3
"""A compact tour of every kernel, for compute-sanitizer where it is available
   (compute-sanitizer --tool memcheck python tools/sanitize_case.py; closed on this pool) and as a
   quick end-to-end exercise otherwise.
Small shapes with awkward sizes (odd widths, chunks shorter than a loop trip, strips hanging over
the image edge), every pixel format, the standalone operators, single-frame and batched detect."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import synth  # noqa: E402

pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
rng = np.random.default_rng(0)
# dense + sparse stages, three formats, streaming and tile kernels, several chunk heights
for shape in [(5, 8), (17, 120), (67, 124), (130, 244), (131, 250), (300, 364)]:
    for dt, ch in ((np.uint8, 0), (np.uint16, 0), (np.uint8, 3)):
        hi = 256 if dt == np.uint8 else 65536
        img = rng.integers(0, hi, shape + ((ch,) if ch else ()), dtype=dt)
        for rows in (0, 4, 58):
            det.set_option("k1_chunk_rows", rows)
            det.stages(img)
        det.set_option("k1_chunk_rows", 0)
        det.refined_saddle_points(img)
    f = rng.random(shape, dtype=np.float32) - 0.5
    for sigma in (1.5, 0.8, 2.6):
        det.gaussian_blur_f32(f, sigma)
    det.hessian_response(f)
# single-frame detect (8 warps per frame) and a small batch (2 warps per frame, both tiers)
img = synth.render_board_numpy(640, 480, seed=3, tag_px=44.0)
tags = det.detect(img)
assert len(tags) == 36, len(tags)
noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
det.detect(noise)
batch = np.stack([img, noise, img[::-1].copy(), img])
res = det.detect_batch(batch)
assert len(res[0]) == 36 and len(res[3]) == 36, [len(r) for r in res]
# device-resident batch + renderer + f32 planes + batched blur
n, W, H = 6, 1280, 1024
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    frames = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
    det.render_boards_device(frames.data_ptr(), n, W, H, 6, 6, 7, stream=s.cuda_stream)
    out = torch.zeros((n, 64 * 9), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    det.detect_batch_device(frames.data_ptr(), n, W, H, pkg.FMT_L8, out.data_ptr(), 64, cnt.data_ptr(),
                            st.data_ptr(), stream=s.cuda_stream)
    src = torch.rand((3, 130, 244), device="cuda")
    dst = torch.empty_like(src)
    det.gaussian_blur_f32_device(src.data_ptr(), 3, 244, 130, 1.5, dst.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
print("tags per frame", cnt.cpu().tolist())
import oracle  # noqa: E402  (test infrastructure: only used to make the planes)
det.detect_planes(oracle.to_luma_f32(img), oracle.to_luma_u8(img))
odd = synth.render_board_numpy(642, 481, seed=5, tag_px=44.0)
det.detect_planes(oracle.to_luma_f32(odd), oracle.to_luma_u8(odd))
det.close()
print("sanitize case done")
