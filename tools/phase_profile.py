#!/usr/bin/env python
"""Developer aid: build the library with RFID_B200_PHASE_PROFILE and print per-phase cycle sums per CTA."""
import os, sys, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import build
lib = os.path.join(ROOT, "gen2_uhf_rfid_reader_b200", "librfid_b200_prof.so")
subprocess.check_call(["nvcc"] + build.NVCC_FLAGS + ["-DRFID_B200_PHASE_PROFILE", "-o", lib, os.path.join(build.CSRC, "rfid_b200.cu")])
build.LIB = lib
import gen2_uhf_rfid_reader_b200.capi as capi
capi.LIB = lib
capi.build_library = lambda: lib
from gen2_uhf_rfid_reader_b200 import synth, abi
dev = torch.device("cuda:0")
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rx = capi.Gen2Rx()
cap = synth.make_capture(nseg, seed=3, device=dev)
segs = capi.segments_to_device(cap["segments"], dev)
tap = torch.zeros((nseg * 2, rx.len_epc), dtype=torch.complex64, device=dev)
for it in range(3):
    rx.decode_capture(cap["iq"], segs, 2)
torch.cuda.synchronize()
rx.set_window_tap(tap)
rx.decode_capture(cap["iq"], segs, 2)
torch.cuda.synchronize()
t = tap.view(torch.int64).cpu().numpy().reshape(-1)[: nseg * 24].reshape(nseg, 24)
names = {0: ["ctl:loop", "ctl:wait_chain", "ctl:flags", "ctl:fsm+e", "ctl:free"],
         8: ["wrk:loop", "wrk:wait_tma", "wrk:blocksum", "wrk:bar1", "wrk:wait_free", "wrk:y+abs", "wrk:bar2", "wrk:d,e"],
         16: ["chn:loop", "chn:wait_full", "chn:wait_elist", "chn:chain+flags"]}
for base, nm in names.items():
    tot = 0
    for i, x in enumerate(nm):
        v = t[:, base + i].mean()
        tot += v
        print("  %-16s %9.0f  per tile %7.0f" % (x, v, v / 27))
    print("  %-16s %9.0f" % ("= total", tot))
print("role end times (cycles since CTA start): chain %.0f worker %.0f control %.0f decoder %.0f" % tuple(t[:, 20 + i].mean() for i in range(4)))
