#!/usr/bin/env python
"""Developer aid: build the library with RFID_B200_PHASE_PROFILE and print the timeline (cycles since CTA start) of the
pack kernel's warps A / B / C of segment 0 and of the chain warp of CTA 0 (GPU box)."""
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
raw_log = tap.view(torch.int64).cpu().numpy().reshape(-1)
log = raw_log[: 320 * 8].reshape(320, 8)
ncta = (nseg + 6) // 7
ct = raw_log[512 * 8: 512 * 8 + 2 * ncta].reshape(ncta, 2).astype(np.float64)
t0 = ct[:, 0].min()
st, en = (ct[:, 0] - t0) / 1000.0, (ct[:, 1] - t0) / 1000.0
print("CTAs %d: start us min %.2f max %.2f | last decode done us: min %.2f p50 %.2f p90 %.2f max %.2f | CTA 0: %.2f .. %.2f" %
      (ncta, st.min(), st.max(), en.min(), np.percentile(en, 50), np.percentile(en, 90), en.max(), st[0], en[0]))
order = np.argsort(en)
print("slowest CTAs:", [(int(i), round(float(en[i]), 2)) for i in order[-6:]], "fastest:", [(int(i), round(float(en[i]), 2)) for i in order[:4]])
A, B, CH, C = log[0:64], log[64:128], log[128:192], log[192:256]
steps = 14 + 2
print("B sub-phases: tile | avg ready  pre-masks  masks  fsm  rebuild  P3done")
for i in range(steps):
    print("%3d | %7d %7d %7d %7d %7d %7d" % (i, B[i][1], B[i][4], B[i][5], B[i][6], B[i][7], B[i][2]))
print("tile | A0: start blocksums |y|stored P1end | A1: start P1end | B: start avg-ready P3done | chain: start inputs-ready done")
A1 = log[256:320] if log.shape[0] >= 320 else A
for i in range(steps):
    print("%3d | %7d %7d %7d %7d | %7d %7d | %7d %7d %7d | %7d %7d %7d" %
          (i, A[i][0], A[i][1], A[i][2], A[i][3], A1[i][0], A1[i][3], B[i][0], B[i][1], B[i][2], CH[i][0], CH[i][1], CH[i][2]))
for k in range(2):
    print("window %d: close seen %d, dc ready %d, decoded %d" % (k, C[k][0], C[k][1], C[k][2]))
for name, d in (("EPC", C[60]), ("RN16", C[61])):
    print("%s decode of segment 0 (cycles): head+sync %d, period search %d, bit samples %d, bits+crc %d, total %d" %
          (name, d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[3], d[4] - d[0]))
