#!/usr/bin/env python3
"""bench.py -- frames/s of the aprilgrid detect path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (TagDetector::detect semantics, every stage) over one
batch of synthetic 1280x1024 u8 gray frames of rendered 6x6 T36H11 AprilGrid boards.
Per-GPU batch is fixed (weak scaling); frames shard image-wise across ranks with no
data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload detect|dense] [--batch B]

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = through
ag_detect_batch with pinned HOST buffers (H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm (the oracle port, all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

W, H = 1280, 1024
METRIC = "frames/sec @1280x1024 gray"
UNIT = "frames/s"
BYTES_PER_PX_DETECT = 17  # K1 9 B/px + K2 8 B/px
BYTES_PER_PX_K1 = 9       # 1 in + 4 blur out + 4 response out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None
        self.t_begin = self.t_end = None  # host clock window of the timed region

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        inside = [r for t, r in self.rows
                  if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.2)]
        for r in (inside or [r for _, r in self.rows[-3:]]):
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and so the first-touch placement of its pinned staging
    buffers) to the NUMA node its GPU hangs off: with one rank per GPU the host-to-device traffic
    of all ranks otherwise funnels through whatever node the ranks happened to start on."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_reference(args):
    """The reference's CPU path (oracle port of aprilgrid-rs detect), frame-parallel on all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    oracle = entry.load_oracle()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    cores = os.cpu_count() or 1
    per_step = max(cores * 2, 16)
    base = synth.fixture_like_frames(8, W, H, seed=100)
    frames = np.ascontiguousarray(np.concatenate([base] * ((per_step + 7) // 8))[:per_step])
    for _ in range(max(args.warmup, 1)):
        oracle.detect_batch(frames[:cores], threads=cores)
    t0 = time.perf_counter()
    n_tags = 0
    for _ in range(args.steps):
        res = oracle.detect_batch(frames, threads=cores)
        n_tags += sum(len(r) for r in res)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "detect_1280x1024_t36h11_6x6", "frames_per_step": per_step,
                   "image": [W, H], "format": "L8"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d rendered 1280x1024 board frames per step, %d steps, one "
                                   "single-threaded detect per frame on %d host threads"
                                   % (per_step, args.steps, cores)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tags_per_frame": n_tags / max(1, args.steps * per_step),
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="detect", choices=["detect", "dense"])
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default 1024; dense 256)")
    ap.add_argument("--chunk", type=int, default=0, help="override pipeline chunk_frames")
    ap.add_argument("--lattice", type=int, default=0, help="override board_lattice (16/32/64)")
    ap.add_argument("--board-warps", type=int, default=-1, help="override board_warps (0 auto, 1/2/4/8)")
    ap.add_argument("--sync-calls", action="store_true",
                    help="order every step's results on the stream before the next step starts "
                         "(default: steps stream through the pipeline, one wait at the end)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sync", action="store_true",
                    help="e2e with synchronous ag_detect_batch calls (default: streaming, two calls in flight)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="extra ag_set_option settings (experiments), e.g. --opt dense_variant=2")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the
        # single JSON line by pointing fd 1 at stderr until the communicator exists.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    pkg = entry.load_package()
    det = pkg.TagDetector(pkg.TagFamily.T36H11, None, device=local)
    if args.chunk:
        det.set_option("chunk_frames", args.chunk)
    if args.lattice:
        det.set_option("board_lattice", args.lattice)
    if args.board_warps >= 0:
        det.set_option("board_warps", args.board_warps)
    for kv in args.opt:
        key, _, val = kv.partition("=")
        det.set_option(key, int(val))
    B = args.batch or (1024 if args.workload == "detect" else 256)
    cap = 64
    stream = torch.cuda.Stream()  # a real (non-default) stream: the kernels and the events share it
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    frames = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    det.render_boards_device(frames.data_ptr(), B, W, H, 6, 6, 1000 + 7919 * rank, stream=sp)
    # two sets of result buffers: with streaming calls, step k+1 starts while step k's board
    # search is still running, so consecutive steps must not share output buffers
    d_tags = [torch.zeros((B, cap * 9), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_cnt = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in range(2)]
    d_status = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in range(2)]
    streaming = args.workload == "detect" and not args.sync_calls
    if streaming:
        det.set_option("device_async", 1)
    step_no = [0]

    def step_device():
        if args.workload == "detect":
            k = step_no[0] & 1
            step_no[0] += 1
            det.detect_batch_device(frames.data_ptr(), B, W, H, pkg.FMT_L8, d_tags[k].data_ptr(), cap,
                                    d_cnt[k].data_ptr(), d_status[k].data_ptr(), stream=sp)
        else:
            det.dense_batch_device(frames.data_ptr(), B, W, H, pkg.FMT_L8, stream=sp)

    def drain():
        if streaming:
            det.detect_batch_device_wait(stream=sp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    drain()
    barrier()
    det.stage_times(reset=True)
    det.set_option("profile", 1)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.5)  # nvidia-smi start-up; its samples are filtered to the timed window below
    launches0 = det.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.t_begin = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    drain()  # every step's results are complete before the closing event
    e1.record(stream)
    barrier()
    if sampler:
        sampler.t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = det.launch_count - launches0
    clocks = sampler.finish() if sampler else None
    det.set_option("profile", 0)
    stage = det.stage_times(reset=True)
    from aprilgrid_rs_b200 import shard
    ms_max = shard.max_over_ranks(ms, device="cuda")
    value = world * B * args.steps / (ms_max * 1e-3)
    # K1 (+K2) alone, same frames, nothing else on the GPU: the roofline figure without the board
    # kernels of earlier chunks sharing the SMs (reported next to the in-step figure)
    k1_alone_ms = None
    if args.workload == "detect":
        torch.cuda.synchronize()
        det.stage_times(reset=True)
        det.set_option("profile", 1)
        det.set_option("dense_variant", 3)  # the same K1 instantiation the detect pipeline launches
        for _ in range(3):
            det.dense_batch_device(frames.data_ptr(), B, W, H, pkg.FMT_L8, stream=sp)
        torch.cuda.synchronize()
        det.set_option("dense_variant", 0)
        det.set_option("profile", 0)
        st_alone = det.stage_times(reset=True)
        k1_alone_ms = st_alone["blur_hessian_min"][0] / max(st_alone["blur_hessian_min"][1], 1)
    cnt_host = d_cnt[0].cpu().numpy() if args.workload == "detect" else None
    if cnt_host is not None:
        assert np.array_equal(cnt_host, d_cnt[1].cpu().numpy()), "steps disagree on the same frames"
    if streaming:
        det.set_option("device_async", 0)

    # ---- end to end through the host-buffer C-ABI call ("e2e") -------------------------------
    e2e = None
    if args.workload == "detect" and not args.no_e2e:
        h_frames = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        h_frames.copy_(frames)
        torch.cuda.synchronize()
        hf = h_frames.numpy()
        h_out = torch.zeros((B, cap * 9), dtype=torch.int32).pin_memory().numpy().view(pkg.TAG_DTYPE).reshape(B, cap)
        h_cnt = torch.zeros(B, dtype=torch.int32).pin_memory().numpy()
        h_status = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        # streaming: a second set of output arrays, two calls in flight (the uploads of step i+1
        # overlap the board searches of step i); every step still uploads its frames and
        # delivers its tags to host memory inside the timed region
        h_out2 = torch.zeros((B, cap * 9), dtype=torch.int32).pin_memory().numpy().view(pkg.TAG_DTYPE).reshape(B, cap)
        h_cnt2 = torch.zeros(B, dtype=torch.int32).pin_memory().numpy()
        h_status2 = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        outs = [(h_out, h_cnt, h_status), (h_out2, h_cnt2, h_status2)]
        streaming = not args.e2e_sync
        det.set_option("host_async", 1 if streaming else 0)

        def host_steps(k):
            for i in range(k):
                det.detect_batch_into(hf, *outs[i & 1])
                if streaming:
                    det.detect_batch_wait(1)
            if streaming:
                det.detect_batch_wait(0)

        host_steps(2)
        barrier()
        t0 = time.perf_counter()
        host_steps(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = shard.max_over_ranks(dt, device="cuda")
        det.set_option("host_async", 0)
        assert np.array_equal(h_cnt, cnt_host), "host-path and device-path results differ"
        assert args.steps < 2 or np.array_equal(h_cnt2, cnt_host), "host-path and device-path results differ"
        assert args.steps < 2 or np.array_equal(h_out2, h_out), "streaming host calls disagree"
        e2e = {"value": world * B * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(world * B * W * H),
               "d2h_bytes_per_step": int(world * B * (cap * 36 + 8)),
               "timing": "host wall clock around ag_detect_batch, pinned host frames, max over ranks",
               "calls": "streaming (host_async): 2 calls in flight, ag_detect_batch_wait" if streaming
                        else "synchronous",
               "numa_node_of_rank0": numa_node}

    # final gather of detections to host: counts only (the records are already on each rank's host)
    total_tags = int(cnt_host.sum()) if cnt_host is not None else 0
    if world > 1 and cnt_host is not None:
        tg = torch.tensor([total_tags], dtype=torch.int64, device="cuda")
        dist.all_reduce(tg)
        total_tags = int(tg.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        k1_ms, k1_n = stage["blur_hessian_min"]
        frames_per_launch = (B * args.steps) / max(k1_n, 1)
        k1_avg_s = (k1_ms / max(k1_n, 1)) * 1e-3
        achieved = BYTES_PER_PX_K1 * W * H * frames_per_launch / max(k1_avg_s, 1e-12) / 1e9
        alg = BYTES_PER_PX_K1 * W * H * frames_per_launch
        in_step = achieved
        if k1_alone_ms:
            # primary figure: K1 timed alone (CUDA events, live in this run, same frames, right after
            # the timed region) against the burst copy bandwidth; inside the streaming step K1
            # shares every SM with the board-search kernels of earlier chunks, so its launch
            # duration there is not a statement about the kernel (kept as in_step_*)
            achieved = alg / (k1_alone_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_blur_hessian_stream (K1: gray->blur->Hessian->min)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src,
                    "traffic": 6.056e9 * frames_per_launch / 512.0,
                    "traffic_source": "profiles/r1g_ncu_full_summary.txt (ncu --set full, 512-frame launch: "
                                      "0.72 GB read + 5.33 GB written = 6.05 GB vs 6.04 GB algorithmic)",
                    "algorithmic_bytes_per_launch": alg,
                    "avg_launch_ms": k1_alone_ms if k1_alone_ms else k1_avg_s * 1e3,
                    "timed": ("K1 alone: CUDA events around each launch, 3 passes over the same frames after "
                              "the timed region" if k1_alone_ms else "CUDA events around each K1 launch"),
                    "in_step_avg_launch_ms": k1_avg_s * 1e3, "in_step_launches_timed": k1_n,
                    "in_step_achieved": in_step, "in_step_frac": in_step / peak,
                    "in_step_note": "same kernel timed inside the streaming step, where it overlaps the "
                                    "board searches of up to 8 earlier chunks",
                    "pipeline_achieved_gbs": BYTES_PER_PX_DETECT * W * H * value / world / 1e9,
                    "pipeline_frac": BYTES_PER_PX_DETECT * W * H * value / world / 1e9 / peak}
        total_stage = sum(v[0] for v in stage.values()) or 1.0
        cpu = None
        if not args.no_cpu:
            oracle = entry.load_oracle()
            cores = os.cpu_count() or 1
            n_cpu = min(B, max(4 * cores, 64))
            sample = frames[:n_cpu].cpu().numpy()
            oracle.detect_batch(sample[:cores], threads=cores)
            t0 = time.perf_counter()
            res = oracle.detect_batch(sample, threads=cores)
            dt = time.perf_counter() - t0
            ok = all(len(r) == int(c) for r, c in zip(res, cnt_host[:n_cpu])) if cnt_host is not None else None
            cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d frames of rank 0's batch, oracle detect (C++ port of the reference), "
                             "one frame per thread on %d threads" % (n_cpu, cores),
                   "tag_counts_equal_gpu": ok}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("detect_1280x1024_t36h11_6x6_batch%d" % B) if args.workload == "detect"
                       else ("dense_blur_hessian_threshold_1280x1024_batch%d" % B),
                       "frames_per_gpu_per_step": B, "image": [W, H], "format": "L8",
                       "parallelism": "frames sharded image-wise, %d rank(s), no collective on the data path" % world,
                       "l2": "inputs (%.2f GB per step) larger than L2" % (B * W * H / 1e9),
                       "calls": "streaming (ag_detect_batch_device with device_async, one wait after the "
                                "last step)" if streaming else "one synchronising call per step"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu,
            "stage_share": {k: v[0] / total_stage for k, v in stage.items()},
            "stage_ms_per_step": {k: v[0] / args.steps for k, v in stage.items()},
            "tags_per_frame": total_tags / float(world * B) if cnt_host is not None else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    det.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
