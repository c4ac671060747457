#!/usr/bin/env python
"""Headline benchmark: images/sec end-to-end (letterbox -> NMS), yolov5s, batch 32, 640x640, fp16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over one batch of 32 synthetic uint8 640x640 images per GPU
(weak scaling: every rank processes its own batch; weights replicated; for N > 1 the padded detections of all
ranks are all-gathered over NCCL inside the step so rank 0 holds a single output list).

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with the uint8 inputs already resident in HBM
(CUDA events around exactly K steps, barrier + synchronize on both sides, max over ranks); `e2e` = same
metric through the public API (`model.predict`) from pinned HOST tensors with the H2D copies and the D2H of the
results inside the timed region.  `roofline` is the tensor-core roofline of the convolution launches (live CUDA
events around the plan); `cpu_baseline` is the CPU oracle port of the reference path on this box's host cores.

`--impl reference` times the reference algorithm's CPU port (oracle/restate.py: the reference's own PyTorch CPU
operators in fp32) with every host thread, on a bounded sample of the same workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 32
SIZE = 640
MODEL = "yolov5s"
SCORE_THRESH = 0.25
GFLOP_PER_IMAGE = {"yolov5n": 4.468, "yolov5s": 16.434, "yolov5m": 48.872, "yolov5l": 108.994, "yolov5x": 205.448}
CPU_SAMPLE_IMAGES = 8          # images per CPU step (bounded sample of the bs32 workload)


def make_state_dict(model, seed: int = 0):
    """Random-init weights of the named architecture (no checkpoints offline): conv ~ N(0, 2/fan_in), BN
    statistics randomised (so folding is exercised), head weights/biases scaled ("load knob") so that about a
    thousand candidates per image pass score_thresh=0.25 and the NMS does real work (300 detections/image).  Pure function of (key, shape)."""
    import zlib

    sd = {}
    for k, v in model.state_dict().items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) & 0x7FFFFFFF)
        shp = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shp, dtype=torch.int64)
        elif k.endswith(".weight") and len(shp) == 4:
            t = torch.randn(shp, generator=g) * (2.0 / (shp[1] * shp[2] * shp[3])) ** 0.5
            if ".head.head." in k:
                t = t * 10.0   # spread the logits (random images give nearly constant features)
        elif k.endswith("bn.weight") or k.endswith("running_var"):
            t = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("bn.bias") or k.endswith("running_mean"):
            t = torch.randn(shp, generator=g) * 0.1
        else:  # head bias [3*(nc+5)]
            b = torch.randn(shp, generator=g).view(3, -1) * 0.1
            b[:, 4] += -3.0
            b[:, 5:] += -4.5
            t = b.reshape(-1)
        sd[k] = t
    return sd


def make_images(n: int, seed0: int):
    ims = []
    for i in range(n):
        g = torch.Generator().manual_seed(seed0 + i)
        ims.append(torch.randint(0, 256, (3, SIZE, SIZE), generator=g, dtype=torch.uint8))
    return ims


class NvmlClockSampler:
    """SM clock / throttle reasons from NVML, polled by a thread while the timed region runs.  In-process (no
    nvidia-smi start-up inside the timed region: spawning it there showed up as multi-millisecond hiccups in
    individual steps, mean 2.27 ms vs median 1.96 ms)."""

    def __init__(self, index: int):
        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        # LOCAL_RANK indexes the visible devices; honour CUDA_VISIBLE_DEVICES when it lists plain indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        phys = int(ids[index]) if index < len(ids) else index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.sm, self.bits, self.stop_flag, self.t = [], 0, False, None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t is not None:
            self.t.join(timeout=1)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx,
                "reasons": [n for n, b in names if self.bits & b], "samples": len(sm), "source": "nvml"}


def make_clock_sampler(index: int):
    try:
        return NvmlClockSampler(index)
    except Exception:
        return ClockSampler(index)     # nvidia-smi -lms fallback


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (fallback when NVML is unusable)."""

    QS = (("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
           "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"),
          ("clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown,"
           "clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            q = self.QS[0]
            for cand in self.QS:   # field names differ between driver generations
                probe = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={cand}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=20)
                if probe.returncode == 0 and len(probe.stdout.strip().split(",")) >= 7:
                    q = cand
                    break
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def conv_traffic():
    """DRAM bytes of all convolution launches of one step, from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_step"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


def cpu_port_images_per_s(sd, steps: int, warmup: int):
    """The reference path restated with the reference's own CPU operators (oracle/restate.py), all host threads."""
    from oracle import restate as R

    ims = make_images(CPU_SAMPLE_IMAGES, 1234)
    # "all the host threads it can use": pick the fastest thread count (oversubscribing a many-core host
    # makes oneDNN/OpenMP convolutions of an 8-image batch much slower than using a few dozen threads)
    ncpu = os.cpu_count() or 1
    best, best_t = ncpu, None
    for nt in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(nt)
        R.detect(sd, ims[:2], score_thresh=SCORE_THRESH)
        t0 = time.perf_counter()
        R.detect(sd, ims[:2], score_thresh=SCORE_THRESH)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        R.detect(sd, ims, score_thresh=SCORE_THRESH)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return CPU_SAMPLE_IMAGES * len(times) / total, times


def run_reference(args, rank, world):
    if rank != 0:
        return
    from yolort_b200.models import yolov5s

    model = yolov5s(score_thresh=SCORE_THRESH)
    sd = make_state_dict(model)
    ips, times = cpu_port_images_per_s(sd, args.steps, args.warmup)
    ms = 1e3 * sum(times) / len(times)
    cores = torch.get_num_threads()
    sample = f"{CPU_SAMPLE_IMAGES} of the {BATCH} images per step ({args.steps} steps after {args.warmup} warm-up), fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": "images/sec end-to-end (letterbox->NMS)", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} batch {BATCH} {SIZE}x{SIZE} uint8 -> detections, score_thresh {SCORE_THRESH}",
                   "sample": sample},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precondition", type=int, default=150,
                    help="untimed steps before the warm-up (clock ramp on a cold box); 0 under a profiler")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    from yolort_b200.models import yolov5s

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = yolov5s(score_thresh=SCORE_THRESH).eval()
    sd = make_state_dict(model)
    model.load_state_dict(sd)
    model = model.to(dev)

    # inputs: NBUF distinct batches (NBUF * 39 MB > the 126 MB L2) so no step finds its input in L2
    NBUF = 4
    host_batches = [torch.stack(make_images(BATCH, 1234 + 1000 * b + 100000 * rank)).pin_memory() for b in range(NBUF)]
    dev_batches = [hb.to(dev) for hb in host_batches]
    D = model.model.post_process.detections_per_img

    def step_device(i):
        b = dev_batches[i % NBUF]
        out = model.forward_padded([b[j] for j in range(BATCH)])
        if world > 1:
            boxes, scores, labels, counts, _ = out
            packed = torch.cat([boxes, scores.unsqueeze(-1), labels.to(torch.float32).unsqueeze(-1)], dim=-1)
            gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, packed)
            gcounts = torch.empty((world, BATCH), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gcounts, counts)
        return out

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident leg ------------------------------------------------------------------------
    plan = model.model.get_plan(BATCH, SIZE, SIZE)
    run_attr = "run_fused" if plan.fused_post is not None else "run"
    native_plan = plan.plan_fused if plan.fused_post is not None else plan.plan
    # Untimed preconditioning before the W warm-up steps: a fresh box idles at ~120 MHz and the first launches also
    # pay module loading; W=3..5 steps are ~10 ms, less than the clock ramp.  --precondition steps (default 150, ~0.3 s).
    # A fixed count (not a time) so that every rank issues the same collectives.
    for i in range(args.precondition):
        step_device(i)
    sync_all()
    for i in range(args.warmup):
        out = step_device(i)
    sync_all()
    sampler = make_clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    plan_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    orig_run = getattr(plan, run_attr)

    step_idx = [0]

    def timed_plan_run(*a_, **k_):
        a, b = plan_ev[step_idx[0]]
        a.record()
        orig_run(*a_, **k_)
        b.record()

    setattr(plan, run_attr, timed_plan_run)
    sync_all()
    ev[0].record()
    for i in range(args.steps):
        step_idx[0] = i
        out = step_device(i)
        ev[i + 1].record()
    sync_all()
    setattr(plan, run_attr, orig_run)
    clocks = None
    if rank == 0:
        try:
            clocks = sampler.stop()
        except Exception as exc:
            clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"sampler failed: {type(exc).__name__}"], "samples": 0}
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    plan_ms = sum(a.elapsed_time(b) for a, b in plan_ev) / args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * BATCH * args.steps / (total_ms / 1e3)
    boxes, scores, labels, counts, status = out
    st = status.cpu().tolist()
    cand_per_img = st[0] / BATCH
    det_per_img = float(counts.float().mean().item())

    # ---- end-to-end leg through the public API, host tensors -----------------------------------------------
    host_lists = [[hb[j] for j in range(BATCH)] for hb in host_batches]
    for i in range(3):
        model.predict(host_lists[i % NBUF])
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h = 0
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        dets = model.predict(host_lists[i % NBUF])          # H2D of 32 uint8 images inside
        # D2H of the results: one copy per field (scores / labels / boxes of the whole batch)
        host_out = {k: torch.cat([d[k] for d in dets]).cpu() for k in ("scores", "labels", "boxes")}
        if i == 0:
            d2h = sum(v.numel() * v.element_size() for v in host_out.values()) + 4 * BATCH + 32
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_value = world * BATCH * args.steps / (float(t.item()) / 1e3)
    d2h_sync = int(d2h)

    # ---- end-to-end leg, throughput API: YOLOv5.predict_stream over the same host batches -----------------------
    # Every step's H2D (39.3 MB from pinned host memory) and D2H (padded detections) is inside the timed region; the
    # API overlaps the copy of batch i+1 with the compute of batch i (two batches in flight).
    e2e_value, e2e_api, d2h, stream_err = e2e_sync_value, "YOLOv5.predict(list of host tensors)", d2h_sync, None
    sync_all()
    try:   # no collective inside the try: a rank that fails must not leave the others waiting in a barrier
        for _ in model.predict_stream(host_lists[i % NBUF] for i in range(3)):
            pass
        torch.cuda.synchronize(dev)
        n_out = 0
        t0 = time.perf_counter()
        e0.record()
        for dets in model.predict_stream(host_lists[i % NBUF] for i in range(args.steps)):
            n_out += len(dets)
        e1.record()
        torch.cuda.synchronize(dev)
        assert n_out == BATCH * args.steps
        stream_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    except Exception as exc:   # keep the line: fall back to the synchronous call's number
        stream_err = f"{type(exc).__name__}: {exc}"[:200]
        stream_ms = -1.0
    t = torch.tensor([stream_ms, -stream_ms], dtype=torch.float64, device=dev)   # max over ranks, and "any rank failed"
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if float(t[1].item()) < 0:      # every rank measured it
        e2e_value = world * BATCH * args.steps / (float(t[0].item()) / 1e3)
        e2e_api = "YOLOv5.predict_stream(iterable of host batches), 2 batches in flight"
        d2h = BATCH * D * 6 * 4 + (BATCH + 4) * 8
    elif stream_err is None:
        stream_err = "another rank failed"

    if rank == 0:
        peak, peak_src = measured_peaks()
        flops = BATCH * GFLOP_PER_IMAGE[MODEL] * 1e9
        achieved = flops / (plan_ms / 1e3) / 1e12
        line = {
            "metric": "images/sec end-to-end (letterbox->NMS)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "p50_ms": step_ms[len(step_ms) // 2],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"{MODEL} batch {BATCH}/GPU {SIZE}x{SIZE} uint8 -> detections (configs[1])",
                       "score_thresh": SCORE_THRESH, "nms_thresh": 0.45, "detections_per_img": D,
                       "candidates_per_image": cand_per_img, "detections_per_image": det_per_img,
                       "l2": f"{NBUF} rotating input batches of 39.3 MB (> 126 MB L2 in total); activations (>1 GB/step) evict everything",
                       "parallelism": f"dp{world}" + (" + nccl all_gather of padded detections" if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": BATCH * 3 * SIZE * SIZE,
                    "d2h_bytes_per_step": int(d2h), "api": e2e_api,
                    "sync_call": {"value": e2e_sync_value, "api": "YOLOv5.predict(list of host tensors)",
                                  "d2h_bytes_per_step": d2h_sync},
                    **({"stream_error": stream_err} if stream_err else {})},
            "gpu_launches": (native_plan.n_ops + 1 + 2 + (0 if plan.fused_post is not None else 1)) * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": conv_traffic(), "peak_source": peak_src, "kernel": "conv_umma_kernel (all conv launches of one step)",
                         "plan_ms": plan_ms, "launches_per_step": native_plan.n_ops},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            ips, times = cpu_port_images_per_s(sd, 2, 1)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{CPU_SAMPLE_IMAGES} of the {BATCH} images x 2 steps after 1 warm-up, fp32 oracle port"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
