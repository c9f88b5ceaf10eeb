"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "aprilgrid_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"AG_API\s+[^;()]*?\b(ag_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["ag_create", "ag_destroy", "ag_detect", "ag_detect_batch", "ag_detect_batch_device",
                 "ag_refined_saddle_points", "ag_gaussian_blur_f32", "ag_gaussian_blur_f32_device", "ag_hessian_response",
                 "ag_family_from_str", "ag_stage_blur", "ag_stage_labels", "ag_stage_saddles"]:
        assert must in syms


def test_streaming_entry_points_are_declared():
    syms = declared_symbols()
    assert "ag_detect_batch_wait" in syms and "ag_detect_batch_device_wait" in syms


def test_cpp_mirror_header_compiles(tmp_path):
    """The C++ host-side mirror of the reference API (cpp/aprilgrid_b200.hpp) is header-only: it must
    at least compile against the C header."""
    src = tmp_path / "use.cpp"
    src.write_text('#include "aprilgrid_b200.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "aprilgrid-rs_b200", "cpp"), str(src)])


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(pkg.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export: " + s
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH]).decode()
    exported = set(re.findall(r" T (ag_[a-z0-9_]+)", out))
    assert set(declared_symbols()) <= exported


def test_library_contains_sm100a_code_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(pkg):
    """On a box without a CUDA device ag_create must fail loudly, never fall back."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as ei:
        pkg.TagDetector(pkg.TagFamily.T36H11)
    assert "no CUDA device" in str(ei.value) or "CPU" in str(ei.value)


def test_default_params_and_family_info(pkg):
    p = pkg._Params()
    pkg.lib().ag_default_params(C.byref(p))
    assert (round(p.tag_spacing_ratio, 6), p.min_saddle_angle, p.max_saddle_angle, p.max_num_of_boards) == \
        (0.3, 30.0, 60.0, 2)  # src/detector.rs:33-40
    for fam, want in [(pkg.TagFamily.T16H5, (4, 2, 1, 30)), (pkg.TagFamily.T25H7, (5, 2, 2, 242)),
                      (pkg.TagFamily.T25H9, (5, 2, 2, 35)), (pkg.TagFamily.T36H11, (6, 2, 3, 587)),
                      (pkg.TagFamily.T36H11B1, (6, 1, 3, 587))]:
        i = pkg.family_info(fam)
        assert (i["edge"], i["border"], i["hamming"], len(i["codes"])) == want


def test_product_does_not_reference_the_oracle():
    """The shipped sources must not include, link or import anything under oracle/."""
    pkg_dir = os.path.join(ROOT, "aprilgrid-rs_b200")
    for dp, _, files in os.walk(pkg_dir):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp", ".hpp", ".rs")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "import oracle" not in txt and "oracle/" not in txt, \
                    os.path.join(dp, f)


def test_board_kernel_shared_memory_budget(pkg):
    """The residency the design counts on: the first launch of a split batch (frames
    of at most 320 saddles, two warps per frame) fits SEVEN blocks per SM (228 KB per SM, 1 KB
    reserved per block) for every image size, every single-frame configuration the API may choose
    fits a block's 227 KB, and the batch launch of the 4096 tier (no grid-ordered positions) fits
    two blocks per SM."""
    import ctypes as C
    lib = pkg.lib()

    def layout(max_saddles, warps, tier, gpos=1, active_cap=0, lattice=64):
        out = (C.c_longlong * 4)()
        assert lib.ag_test_board_layout(max_saddles, lattice, warps, tier, gpos, active_cap, out) == 0
        return [int(v) for v in out]

    sm_bytes, reserve, block_max = 228 * 1024, 1024, 227 * 1024
    for max_saddles in (2048, 5120, 13056):  # automatic capacities of 1280x1024, 2048x1536, 3840x2160 frames
        smem, _, warps, tier = layout(max_saddles, 2, 320, active_cap=320)
        assert warps == 2 and tier == 320
        assert 7 * (smem + reserve) <= sm_bytes, (max_saddles, smem)
    for max_saddles in (2048, 5120):
        for t in (320, 512, 1024):
            assert layout(max_saddles, 16, t)[0] <= block_max, (max_saddles, t)
    big = 13056
    assert 2 * (layout(big, 2, 4096, gpos=0)[0] + reserve) <= sm_bytes
    assert layout(big, 8, 4096)[0] <= block_max
    assert layout(big, 16, 4096)[0] > block_max  # the API halves the warps here (ensure_board_slot)
