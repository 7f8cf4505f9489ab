#!/usr/bin/env python
"""bench.py -- MSamples/s of raw I/Q through matched filter -> gate -> tag_decoder.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic 40 kHz-BLF FM0 I/Q @ 2 MS/s, 1000 queries (= 1000
inventory-round segments of 16,960 raw samples), 1 tag, per GPU.  One *step* = one pass of the hot path over
one such capture.  `value` is timed with the captures already resident in HBM; `e2e` goes through the
host-pointer C-ABI call with pinned host buffers (H2D of the capture and D2H of the records inside the
timed region).  Weak scaling: every rank decodes its own 1000-round shard of the global segment table; the
one collective is a single all-gather of the decoded records of all K steps at the end of the timed region.

--impl reference times the reference's own CPU implementation (oracle/_ref: its blocks compiled unchanged)
on all host cores over the same workload.
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

N_ROUNDS = 1000
SEG_US = 8480.0
ADC_RATE = 2_000_000
NBUF = 4          # distinct captures cycled through so every step reads data that is not in L2
MAX_WINDOWS = 2
METRIC = "MSamples/s I/Q through gate->tag_decoder"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML from a background thread (every ~0.5 ms) so that even a
    millisecond-long timed region gets samples; `window(t0, t1)` summarises the samples taken inside it."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th, self.ok = index, [], False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        if not self.ok:
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.0004)

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=1)

    def window(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        pad = 0.0
        while len(inside) < 3 and pad < 0.2:   # very short region: widen symmetrically, say so
            pad += 0.005
            inside = [r for r in self.rows if t0 - pad <= r[0] <= t1 + pad]
        reasons = set()
        for r in inside:
            for bit, nm in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(reasons), "samples": len(inside), "window_pad_ms": round(pad * 1e3, 1)}


# ----------------------------------------------------------------------------------------- reference arm
_G = {}


def _ref_init(kind):
    from oracle import pyoracle, refflow
    _G["flow"] = refflow.RefFlow(0) if kind == "reference" else pyoracle.Oracle()
    _G["kind"] = kind


def _ref_worker(args):
    seg_bytes, passes = args
    from gen2_uhf_rfid_reader_b200 import abi
    segs = np.frombuffer(seg_bytes, dtype=abi.SEGMENT_DTYPE)
    iq = _G["iq"]          # inherited through fork (copy-on-write, never written)
    t, n = 0.0, 0
    for _ in range(passes):
        if _G["kind"] == "reference":
            _, counts, secs = _G["flow"].run_segments(iq, segs, max_per_seg=MAX_WINDOWS, want_records=True)
        else:
            _, counts, secs = _G["flow"].decode_segments(iq, segs, max_per_seg=MAX_WINDOWS)
        t += secs
        n += int(counts.sum())
    return t, n


def host_cores():
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota (a container that sees
    128 logical CPUs but is granted 16 CPUs of time is a 16-core host for this purpose -- more runnable
    processes than that only get throttled)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = int(q) / int(p)
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota:
        n = min(n, max(1, int(quota + 0.5)))
    return max(1, n)


def cpu_reference_run(iq_np, segs, steps, warmup, passes_per_step=1):
    """Reference CPU implementation on all host cores: one process per core over disjoint segment ranges
    (the reference keeps its state in a process global, include/rfid/global_vars.h:146).  Timed by wall
    clock around each step (all processes working), excluding data generation and file I/O."""
    import multiprocessing as mp
    from oracle import refflow
    kind = "reference" if refflow.ref_available(0) else "port"
    cores = host_cores()
    nseg = segs.size
    _G["iq"] = np.ascontiguousarray(iq_np)
    jobs = []
    for c in range(cores):
        b, e = c * nseg // cores, (c + 1) * nseg // cores
        jobs.append((segs[b:e].copy().tobytes(), passes_per_step))
    ctx = mp.get_context("fork")
    times, windows = [], 0
    with ctx.Pool(cores, initializer=_ref_init, initargs=(kind,)) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = pool.map(_ref_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
                windows = sum(o[1] for o in out)
    _G.pop("iq", None)
    n_samples = float(segs["length"].astype(np.float64).sum()) * passes_per_step
    total = sum(times)
    return {"kind": kind, "cores": cores, "ms_per_step": 1e3 * total / max(1, len(times)),
            "value": n_samples * len(times) / total / 1e6, "windows_last_step": windows,
            "sample": "cfg2 capture: %d segments (%.1f M samples) split over %d processes (= usable CPUs: affinity and "
                      "cgroup quota; os.cpu_count() = %d), %d pass(es) per step, %d steps"
                      % (nseg, n_samples / passes_per_step / 1e6, cores, os.cpu_count() or 1, passes_per_step, len(times))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rounds", type=int, default=N_ROUNDS, help="inventory rounds (segments) per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--generator", default="torch", choices=["torch", "native"],
                    help="workload generator of our arm: the torch model (same samples as the CPU reference arm) or the "
                         "library's CUDA closed-loop slot simulator (rfid_b200_sim_capture)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from gen2_uhf_rfid_reader_b200 import abi, synth
    seg_len = int(round(SEG_US * ADC_RATE / 1e6))
    config = {"workload": "cfg2: synthetic 40kHz-BLF FM0 I/Q @2Msps, %d queries (inventory rounds) x %d raw samples, 1 tag, per GPU"
                          % (args.rounds, seg_len),
              "rounds_per_gpu": args.rounds, "segment_samples": seg_len, "fixed_q": 0,
              "l2": "inputs larger than L2: %d distinct captures of %.0f MB cycled" % (NBUF, args.rounds * seg_len * 8 / 1e6),
              "parallelism": "segments sharded over %d GPU(s), one all-gather of all decoded records at the end of the timed region" % world,
              "generator": args.generator}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cap = synth.make_capture(args.rounds, seed=1234, device="cpu")
        # each step = 8 passes over the 1000-round capture (bounded sample: ~0.05-0.1 s per step on a 16-CPU host)
        r = cpu_reference_run(cap["iq"].numpy(), cap["segments"], args.steps, args.warmup, passes_per_step=8)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MSamples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from gen2_uhf_rfid_reader_b200 import capi, shard
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_global = args.rounds * world
    first = args.rounds * rank

    rx = capi.Gen2Rx(device=local_rank)
    caps, truths, seg_dev = [], [], None
    for b in range(NBUF):
        if args.generator == "native":
            sim = capi.default_sim(seed=1234 + 17 * b, segment_us=SEG_US)
            cap = rx.sim_capture(sim, args.rounds, first_segment=first, device=dev)
            tr = cap["truth"].cpu().numpy().view(abi.SIM_TRUTH_DTYPE).reshape(-1)
            caps.append(cap["iq"])
            truths.append({"rn16": tr["acked_rn16"].astype(np.int64), "epc": tr["epc"]})
            if seg_dev is None:
                seg_dev = cap["segs"]
                segs_np = cap["segs"].cpu().numpy().view(abi.SEGMENT_DTYPE).reshape(-1)
        else:
            cap = synth.make_capture(args.rounds, seed=1234 + 17 * b, first_segment=first, device=dev)
            caps.append(cap["iq"])
            truths.append(cap["truth"])
            if seg_dev is None:
                segs_np = cap["segments"]
                seg_dev = capi.segments_to_device(segs_np, dev)
    n_raw = caps[0].numel()
    # every step keeps its records on the device; ONE all-gather of all of them closes the timed region
    # (north star: "a single NCCL gather of decoded EPCs at the end")
    nslots = max(args.steps, args.warmup)
    results_all = torch.zeros((nslots, args.rounds * MAX_WINDOWS, 64), dtype=torch.uint8, device=dev)
    counts_all = torch.zeros((nslots, args.rounds), dtype=torch.int32, device=dev)
    g_res = torch.empty((world,) + tuple(results_all.shape), dtype=torch.uint8, device=dev) if world > 1 else None
    g_cnt = torch.empty((world,) + tuple(counts_all.shape), dtype=torch.int32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream(dev)
    launches = 0

    def step(i, slot):
        rx.decode_capture(caps[i % NBUF], seg_dev, MAX_WINDOWS, results_all[slot], counts_all[slot], stream)

    def gather():
        if world > 1:
            dist.all_gather_into_tensor(g_res, results_all)
            dist.all_gather_into_tensor(g_cnt, counts_all)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i, i)
    gather()
    barrier()
    # correctness gate on the warm-up output: every round must decode its tag's EPC with a valid CRC
    recs, cnt = capi.results_to_numpy(results_all[args.warmup - 1], counts_all[args.warmup - 1], MAX_WINDOWS)
    epc_ok = int((recs[:, 1]["crc_ok"] == 1).sum())
    truth = truths[(args.warmup - 1) % NBUF]
    rn_ok = int((recs[:, 0]["tag_id"] == truth["rn16"]).sum())
    epc_match = int((recs[:, 1]["bits"] == truth["epc"]).all(axis=1).sum())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rx.enable_kernel_timing(True)
    rx.kernel_time(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, i)
        launches += rx.last_launch_count()
    gather()
    ev1.record(stream)
    barrier()
    t_host1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    k_ms, k_n = rx.kernel_time(reset=True)
    rx.enable_kernel_timing(False)
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.window(t_host0, t_host1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = args.steps * n_raw * world / (ms_max * 1e-3) / 1e6

    # ------------------------------------------------------------------ end-to-end: host buffers through the C-ABI
    h_iq = [torch.empty(n_raw, dtype=torch.complex64).pin_memory() for _ in range(2)]
    for b in range(2):
        h_iq[b].copy_(caps[b])
    h_res = torch.zeros((args.rounds * MAX_WINDOWS, 64), dtype=torch.uint8).pin_memory()
    h_cnt = torch.zeros(args.rounds, dtype=torch.int32).pin_memory()
    h_segs = torch.from_numpy(np.ascontiguousarray(segs_np).view(np.uint8).copy()).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(i):
        rx.decode_capture_host_ptr(h_iq[i % 2].data_ptr(), n_raw, h_segs.data_ptr(), args.rounds, MAX_WINDOWS,
                                   h_res.data_ptr(), h_cnt.data_ptr())

    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * n_raw * world / float(te.item()) / 1e6
    e2e_recs = h_res.numpy().reshape(-1).view(abi.RESULT_DTYPE).reshape(args.rounds, MAX_WINDOWS)
    e2e_ok = int((e2e_recs[:, 1]["crc_ok"] == 1).sum())

    if rank == 0:
        peak, peak_kind = hbm_peak()
        k_avg_ms = k_ms / max(1, k_n)
        achieved = 8.0 * n_raw / (k_avg_ms * 1e-3) / 1e9 if k_n else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("rounds") == args.rounds:
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            except Exception:
                pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "kernel": "rx_fused_split_kernel<5,5> (matched filter + gate + tag_decoder, one CTA per segment)",
                    "kernel_ms": k_avg_ms, "kernel_launches_timed": k_n, "peak_kind": peak_kind,
                    "algorithmic_bytes_per_launch": 8.0 * n_raw}
        cpu_b = None
        if not args.no_cpu_baseline:
            cap_cpu = synth.make_capture(args.rounds, seed=1234, device="cpu")
            # bounded sample: ~20-30 s of CPU work = 6 x 40 passes over the 1000-round capture, all usable CPUs
            r = cpu_reference_run(cap_cpu["iq"].numpy(), cap_cpu["segments"], steps=5, warmup=1, passes_per_step=40)
            cpu_b = {"value": r["value"], "unit": "MSamples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        line = {"metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": int(n_raw * 8 + h_segs.numel()),
                        "d2h_bytes_per_step": int(h_res.numel() + h_cnt.numel() * 4), "steps": e2e_steps,
                        "epc_crc_ok": e2e_ok},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_b,
                "parity": {"epc_crc_ok": epc_ok, "rn16_match_truth": rn_ok, "epc_match_truth": epc_match, "rounds": args.rounds}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
