#!/usr/bin/env python
"""bench.py -- DiMP-50 tracked frames/sec on synthetic 288x288 search crops (BASELINE.json configs[1]).

One "step" = one tracked frame of the hot path on every rank: search crop -> ResNet-50 backbone to layer3 ->
clf head (conv3x3 + InstanceL2Norm) -> apply_filter + max2d -> localisation decision (host) -> memory update ->
10 steepest-descent iterations over the full 50-sample memory ("10 SD iters/frame").

  value : frames/s with the crops already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : frames/s through the host-buffer C-ABI frame calls (pinned host crop in, score map + arg-max out;
          H2D/D2H inside the timed region, host-side localisation + sample-weight bookkeeping included)
  --impl reference : the same per-frame path as the CPU restatement of the reference (oracle/, torch-CPU fp32,
          all host threads) -- the reference itself cannot travel to the GPU box (no /root/reference there).

Multi-GPU: one process per GPU, one independent synthetic sequence per rank (weak scaling), no collective
inside the frame loop; a single NCCL all_gather of per-rank timings at the end.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CROP = 288
MEMORY = 50
SD_ITERS = 10
# dram__bytes_read.sum + dram__bytes_write.sum of one sd_kernel launch (n=50, 10 it) from the committed ncu --set full
# capture profiles/r01g_ncu_full_sd_and_conv.txt: the sample memory is read from HBM once per call and stays L2 resident.
SD_DRAM_TRAFFIC_BYTES = 33330944 + 121600
# the same for one sd_tc_kernel launch on the frame engine's pitched sample memory (profiles/r01j_ncu_full_sd_tc.txt)
SD_TC_DRAM_TRAFFIC_BYTES = 36231936 + 275200
POOL = 160          # distinct crops per rank (160 x 995 KB = 159 MB > 126 MB L2: a step's input is never L2 resident)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", type=int, default=0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def localisation_decision(scores, max_val, max_idx, prev_disp, target_cells=(3.1, 3.1), p=None):
    """Host half of DiMP.localize_advanced (pytracking/tracker/dimp/dimp.py:238-303) on the 19x19 map of one scale:
    threshold tests, neighbourhood masking, second maximum, distractor / hard-negative decision tree."""
    p = p or {}
    sz = scores.shape[-1]
    s1 = float(max_val)
    r1, c1 = int(max_idx[0]), int(max_idx[1])
    if s1 < p.get("target_not_found_threshold", -1e9):
        return "not_found", (r1, c1)
    ny, nx = p.get("target_neighborhood_scale", 2.2) * target_cells[0], p.get("target_neighborhood_scale", 2.2) * target_cells[1]
    top, bot = max(round(r1 - ny / 2), 0), min(round(r1 + ny / 2 + 1), sz)
    lef, rig = max(round(c1 - nx / 2), 0), min(round(c1 + nx / 2 + 1), sz)
    masked = scores.copy()
    masked[top:bot, lef:rig] = 0
    col_best = masked.max(axis=0)
    c2 = int(col_best.argmax())
    r2 = int(masked[:, c2].argmax())
    s2 = float(masked[r2, c2])
    ctr = (sz - 1) / 2
    d1 = np.hypot(r1 - ctr - prev_disp[0], c1 - ctr - prev_disp[1])
    d2 = np.hypot(r2 - ctr - prev_disp[0], c2 - ctr - prev_disp[1])
    thr = p.get("dispalcement_scale", 0.8) * np.sqrt(sz * sz) / 2
    if s2 > p.get("distractor_threshold", 0.8) * s1:
        if d2 > thr and d1 < thr:
            return "hard_negative", (r1, c1)
        if d2 < thr and d1 > thr:
            return "hard_negative", (r2, c2)
        return "uncertain", (r1, c1)
    if s2 > p.get("hard_negative_threshold", 0.5) * s1 and s2 > p.get("target_not_found_threshold", -1e9):
        return "hard_negative", (r1, c1)
    return "normal", (r1, c1)


def setup_engine(rank, precision):
    from pytracking_b200 import synth
    from pytracking_b200.frame_engine import DiMPFrameEngine, SampleWeights
    sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
    eng = DiMPFrameEngine(sd, arch="resnet50", filter_size=4, memory_size=MEMORY, max_batch=1, crop_size=CROP,
                          precision=precision)
    # fill the sample memory as DiMP.initialize would (15 init samples) and then to capacity (steady state)
    q = 1000 + rank
    init = synth.make_crop(q, MEMORY, CROP)
    boxes = synth.make_boxes(q + 1, MEMORY)
    for i in range(MEMORY):
        out = eng.backbone.forward(init[i:i + 1].cuda(), want=("classification",))
        eng.memory[i].copy_(out["classification"][0])
    eng.boxes.copy_(boxes.cuda())
    sw = SampleWeights(MEMORY, 15)
    for _ in range(MEMORY - 15):
        sw.step()
    eng.sample_weights.copy_(torch.from_numpy(sw.w).cuda())
    eng.filter.zero_()
    eng.localize_device(init[MEMORY - 1:MEMORY].cuda())          # leaves the clf feature of the last init crop in the state
    eng.update(0, MEMORY - 1, boxes[MEMORY - 1].numpy(), sw.w, MEMORY, SD_ITERS)   # initial model: 10 SD iterations from w = 0
    torch.cuda.synchronize()
    return eng, sw, synth


def run_b200(args, rank, world, local_rank):
    from pytracking_b200 import _lib
    torch.cuda.set_device(local_rank)
    eng, sw, synth = setup_engine(rank, args.precision)
    dev = eng.device
    g = torch.Generator().manual_seed(7000 + rank)
    pool_host = torch.empty(POOL, 3, CROP, CROP, dtype=torch.float32).pin_memory()
    base = synth.make_crop(2000 + rank, 8, CROP)
    for i in range(POOL):   # cheap distinct crops: shifted / re-noised variants of 8 base crops
        pool_host[i] = torch.roll(base[i % 8], shifts=(i * 3) % CROP, dims=2)
    pool_dev = pool_host.to(dev)
    boxes_np = synth.make_boxes(3000 + rank, POOL).numpy()
    K, W = args.steps, args.warmup
    barrier = (lambda: torch.distributed.barrier()) if world > 1 else (lambda: None)

    def frame_device(i):
        eng.localize_device(pool_dev[i % POOL:i % POOL + 1])
        r = sw.step()
        eng.update(0, r, boxes_np[i % POOL], sw.w, MEMORY, SD_ITERS)

    prev = [0.0, 0.0]

    def frame_host(i):
        scores, mv, mi = eng.localize(pool_host[i % POOL:i % POOL + 1])
        flag, (r_, c_) = localisation_decision(scores[0], mv[0], mi[0], prev)
        prev[0], prev[1] = r_ - 9.0, c_ - 9.0
        lr = 0.02 if flag == "hard_negative" else None
        r = sw.step(lr)
        eng.update(0, r, boxes_np[i % POOL], sw.w, MEMORY, SD_ITERS)

    # ---------------- device-resident leg (value) ----------------
    for i in range(W):
        frame_device(i)
    torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.25)
    launches0 = _lib.lib().b200trk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); barrier()
    e0.record()
    for i in range(K):
        frame_device(W + i)
    e1.record()
    torch.cuda.synchronize(); barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = _lib.lib().b200trk_launch_count() - launches0

    # per-kernel event timing of the dominant kernel (SD optimiser) for the roofline entry
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * 20)]
    for j in range(20):
        eng.localize_device(pool_dev[j:j + 1])
        ev[2 * j].record()
        eng.update(0, j % MEMORY, boxes_np[j], sw.w, MEMORY, SD_ITERS)
        ev[2 * j + 1].record()
    torch.cuda.synchronize()
    sd_us = float(np.median([ev[2 * j].elapsed_time(ev[2 * j + 1]) for j in range(20)])) * 1e3
    sd_tc = int(_lib.lib().b200trk_sd_last_kernel())

    # ---------------- host-buffer leg (e2e) ----------------
    for i in range(W):
        frame_host(i)
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    for i in range(K):
        frame_host(W + i)
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.finish() if sampler else None

    # ---------------- gather: max over ranks ----------------
    t = torch.tensor([dev_ms, host_ms, sd_us], device=dev, dtype=torch.float64)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        torch.distributed.all_gather(allt, t)
        t = torch.stack(allt).max(dim=0).values
    dev_ms, host_ms, sd_us = [float(x) for x in t.cpu()]
    if rank != 0:
        return
    peaks = load_peaks()
    # SD-GN algorithmic bytes per iteration = 3*n*663552 + 8*n*1444 + 3*32768 B (n = 50)
    sd_bytes = SD_ITERS * (3 * MEMORY * 663552 + 8 * MEMORY * 1444 + 3 * 32768)
    achieved = sd_bytes / (sd_us * 1e-6) / 1e9
    cpu = cpu_baseline_sample(steps=3, warmup=1)
    out = {
        "metric": "DiMP-50 tracked frames/sec (288x288 synthetic search crops, 10 SD iters/frame)",
        "value": world * K / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (3xTF32 error-compensated tcgen05 convolutions and optimiser sweeps; fp32 CUDA-core correlation)" if args.precision == 0 else "f32",
        "data": "synthetic",
        "config": {"workload": "DiMP-50 single-GPU, synthetic 288x288 crops, 10 SD iters/frame (BASELINE configs[1]); "
                               "one independent sequence per GPU", "backbone": "resnet50->layer3", "memory": MEMORY,
                   "sd_iters": SD_ITERS, "use_iou_net": False,
                   "l2": "inputs > L2: each step reads a different crop from a %d-crop pool (%.0f MB); network weights and the "
                         "50-sample memory are the tracker's steady-state working set" % (POOL, POOL * 3 * CROP * CROP * 4 / 1e6),
                   "precision": args.precision},
        "e2e": {"value": world * K / (host_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 3 * CROP * CROP * 4 + 16 + MEMORY * 4,
                "d2h_bytes_per_step": 19 * 19 * 4 + 4 + 16},
        "gpu_launches": int(launches),
        "roofline": {"kernel": ("sd_tc_kernel<18,0> (tcgen05 sweeps)" if sd_tc else "sd_kernel<18,4,0>") + " (DiMP steepest-descent, n=50, 10 it)",
                     "bound": "hbm", "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": SD_TC_DRAM_TRAFFIC_BYTES if sd_tc else SD_DRAM_TRAFFIC_BYTES,
                     "peak_source": peaks["source"], "us_per_launch": sd_us, "us_per_sd_iteration": sd_us / SD_ITERS},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------
class CpuFrame:
    """The same per-frame path on the host: torch-CPU restatement of the reference (oracle/dimp_oracle.py)."""

    def __init__(self):
        from oracle import dimp_oracle as O
        from pytracking_b200 import synth
        from pytracking_b200.frame_engine import SampleWeights
        self.O, self.synth = O, synth
        self.sd = synth.make_dimp_state_dict("resnet50", seed=0, lut_seed=3)
        self.p = {k[len("classifier.filter_optimizer."):]: v for k, v in self.sd.items() if k.startswith("classifier.filter_optimizer.")}
        self.memory = synth.make_clf_features(11, MEMORY, 512, 18, 18)
        self.boxes = synth.make_boxes(12, MEMORY)
        self.sw = SampleWeights(MEMORY, 15)
        for _ in range(MEMORY - 15):
            self.sw.step()
        self.w = torch.zeros(1, 512, 4, 4)
        self.crops = synth.make_crop(13, 4, CROP)

    def step(self, i):
        O = self.O
        with torch.no_grad():
            im = O.preprocess_image(self.crops[i % 4:i % 4 + 1])
            bf = O.resnet_forward(self.sd, im, "resnet50", output_layers=("layer3",))
            clf = O.clf_head_dimp50(self.sd, bf["layer3"])
            s = O.apply_filter_conv(clf, self.w)
            O.max2d(s[:, 0])
            r = self.sw.step()
            self.memory[r] = clf[0]
            self.w = O.dimp_sd_gn_conv(self.w, self.memory, self.boxes, torch.from_numpy(self.sw.w), self.p, SD_ITERS)


def host_cores():
    """Cores this process may really use: scheduler affinity, capped by the cgroup CPU quota if there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


def pick_cpu_threads(frame):
    """The torch-CPU path does not scale to every core count (tiny grouped convs); time one frame per candidate
    thread count and keep the fastest, so that the baseline is the best the host can do, not an oversubscribed one."""
    avail = host_cores()
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail} or {avail})
    best, best_t = cands[0], None
    for c in cands:
        torch.set_num_threads(c)
        frame.step(0)
        t0 = time.perf_counter()
        frame.step(1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
        if dt > 8.0:          # hopeless at this count; larger counts only get worse
            break
    torch.set_num_threads(best)
    return best


def cpu_baseline_sample(steps, warmup):
    f = CpuFrame()
    pick_cpu_threads(f)
    for i in range(warmup):
        f.step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        f.step(warmup + i)
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d frames of the same DiMP-50 hot path (backbone+head+classify+10 SD it over 50 samples), torch-CPU fp32 "
                      "restatement of the reference (oracle/dimp_oracle.py)" % steps}


def run_reference(args, rank, world):
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    K = min(K, 30)        # bounded sample: ~0.2 s per frame on host cores
    W = min(W, 3)
    f = CpuFrame()
    pick_cpu_threads(f)
    for i in range(W):
        f.step(i)
    t0 = time.perf_counter()
    for i in range(K):
        f.step(W + i)
    dt = time.perf_counter() - t0
    v = K / dt
    out = {
        "impl": "reference",
        "metric": "DiMP-50 tracked frames/sec (288x288 synthetic search crops, 10 SD iters/frame)",
        "value": v, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DiMP-50 single-GPU, synthetic 288x288 crops, 10 SD iters/frame (BASELINE configs[1])",
                   "backbone": "resnet50->layer3", "memory": MEMORY, "sd_iters": SD_ITERS, "use_iou_net": False},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d frames, torch-CPU fp32 restatement of the reference path (the reference tree is not "
                                   "available on the GPU box)" % K},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    run_b200(args, rank, world, local_rank)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
