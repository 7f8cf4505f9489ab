#!/usr/bin/env python
"""Developer aid: build the library with RFID_B200_PHASE_PROFILE and print the per-phase cycle sums of the pack kernel's
tile warps and chain warps (GPU box)."""
import os, sys, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import build
lib = os.path.join(ROOT, "gen2_uhf_rfid_reader_b200", "librfid_b200_prof.so")
if not os.path.exists(lib) or os.environ.get("REBUILD"):
    subprocess.check_call(["nvcc"] + build.NVCC_FLAGS + ["-DRFID_B200_PHASE_PROFILE", "-o", lib, os.path.join(build.CSRC, "rfid_b200.cu")])
os.environ["RFID_B200_LIB"] = lib
import gen2_uhf_rfid_reader_b200.capi as capi
from gen2_uhf_rfid_reader_b200 import synth
dev = torch.device("cuda:0")
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = int(os.environ.get("RFID_B200_PACK_G", "0")) or min(7, -(-nseg // 148))
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
ncta = -(-nseg // G)
flat = tap.view(torch.int64).cpu().numpy().reshape(-1)
t = flat[: (nseg + ncta) * 8].reshape(nseg + ncta, 8)
steplog = flat[(nseg + ncta) * 8: (nseg + ncta) * 8 + 512 * 8].reshape(512, 8)
tile, chain = t[:nseg], t[nseg:]
steps = 27 + 3
for nm, col in (("P1 wait tma", 0), ("P1 compute", 1), ("arrive X + wait Y", 2), ("E", 3), ("P3", 4), ("loop", 5)):
    v = tile[:, col].mean()
    print("tile  %-18s %9.0f  per step %7.0f" % (nm, v, v / steps))
print("tile  total              %9.0f" % tile[:, :6].sum(axis=1).mean())
for nm, col in (("wait X", 0), ("chain", 1), ("arrive Y", 2)):
    v = chain[:, col].mean()
    print("chain %-18s %9.0f  per step %7.0f" % (nm, v, v / steps))
print("chain total              %9.0f" % chain[:, :3].sum(axis=1).mean())

print("per step, tile warp of segment 0:  step | waitTMA P1 wait E P3 loop || chain: waitX chain arrive")
for i in range(steps):
    a, c = steplog[i], steplog[64 + i]
    print("%3d | %6d %6d %6d %6d %6d %6d || %6d %6d %6d" % (i, a[0], a[1], a[2], a[3], a[4], a[5], c[0], c[1], c[2]))

print("warp B of segment 0:  step | tma-issue waitY E P3 pair-wait || E: flush, loop start, loop end (since phase start) | P3: pre-masks, masks, fsm")
for i in range(steps):
    b = steplog[192 + i]; sb = steplog[256 + i]
    print("%3d | %6d %6d %6d %6d %6d || %6d %6d %6d | %6d %6d %6d" % (i, b[0], b[2], b[3], b[4], b[5], sb[0], sb[1], sb[2], sb[3], sb[4], sb[5]))
print("P1 sub-phases (cycles since the TMA wait ended): blocksums | +tma issue | +shuffles | +MF,cabsf,store | +lookback,range | P1 end")
for i in range(0, steps, 3):
    b = steplog[128 + i]
    print("%3d | %6d %6d %6d %6d %6d | %6d" % (i, b[0], b[1], b[2], b[3], b[4], steplog[i][1]))

print("absolute times (cycles since CTA start), segment 0")
print("step | A: P1start P1end pairdone | B: Ydone Edone P3done pairdone | chain: Xdone done")
for i in range(steps):
    a, b, c = steplog[320 + i], steplog[384 + i], steplog[448 + i]
    # A marks: 5 (loop top = after pair), 0 (tma ready), 1 (P1 end);  B marks: 5 (top), 0, 2 (Y done), 3 (E done), 4 (P3 done); chain: 0 (X done), 1 (done), 2
    print("%3d | %7d %7d %7d | %7d %7d %7d %7d | %7d %7d" % (i, a[0], a[1], steplog[320 + i + 1][5] if i + 1 < steps else 0, b[2], b[3], b[4], steplog[384 + i + 1][5] if i + 1 < steps else 0, c[0], c[1]))
