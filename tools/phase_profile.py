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
t = tap.view(torch.int64).cpu().numpy().reshape(-1)[: nseg * 16].reshape(nseg, 16)
seqn = ["tail", "wait_full", "chain", "finalize", "flags", "fsm+e"]
wrk = ["loop", "wait_tma", "blocksum", "bar1", "wait_empty", "y+abs", "bar2", "d+arrive"]
print("sequencer phases (mean cycles per segment over CTAs):")
for i, nme in enumerate(seqn):
    print("  %-10s %9.0f  per tile %7.0f" % (nme, t[:, i].mean(), t[:, i].mean() / 27))
print("  total      %9.0f" % t[:, :6].sum(axis=1).mean())
print("worker phases:")
for i, nme in enumerate(wrk):
    print("  %-10s %9.0f  per tile %7.0f" % (nme, t[:, 8 + i].mean(), t[:, 8 + i].mean() / 27))
print("  total      %9.0f" % t[:, 8:16].sum(axis=1).mean())
