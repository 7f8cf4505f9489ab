import os, sys, subprocess
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import build
lib = os.path.join(ROOT, "gen2_uhf_rfid_reader_b200", "librfid_b200_prof.so")
subprocess.check_call(["nvcc"] + build.NVCC_FLAGS + ["-DRFID_B200_PHASE_PROFILE", "-o", lib, os.path.join(build.CSRC, "rfid_b200.cu")])
os.environ["RFID_B200_LIB"] = lib
import gen2_uhf_rfid_reader_b200.capi as capi
from gen2_uhf_rfid_reader_b200 import synth
dev = torch.device("cuda:0")
nseg = 1000
rx = capi.Gen2Rx()
cap = synth.make_capture(nseg, seed=3, device=dev)
segs = capi.segments_to_device(cap["segments"], dev)
tap = torch.zeros((nseg * 2, rx.len_epc), dtype=torch.complex64, device=dev)
rx.decode_capture(cap["iq"], segs, 2); torch.cuda.synchronize()
rx.set_window_tap(tap)
rx.decode_capture(cap["iq"], segs, 2); torch.cuda.synchronize()
t = tap.view(torch.int64).cpu().numpy().reshape(-1)
m = t[nseg * 24: nseg * 24 + nseg * 4].reshape(nseg, 4)
smid = m >> 16; wid = m & 0xFFFF
names = ["chain", "control", "worker", "decoder"]
print("warpid % 4 histogram per role:")
for r in range(4):
    print("  %-8s" % names[r], np.bincount(wid[:, r] % 4, minlength=4))
# per SM: how many chains per SMSP
cnt = {}
for c in range(nseg):
    cnt.setdefault(int(smid[c, 0]), []).append(int(wid[c, 0] % 4))
dist = np.array([np.bincount(v, minlength=4) for v in cnt.values()])
print("chains per sub-partition (per SM), mean over SMs of sorted counts:", np.sort(dist, axis=1).mean(axis=0))
print("CTAs per SM:", np.bincount([len(v) for v in cnt.values()]))
print("example SM:", smid[:, 0][:8], [sorted((int(wid[c, r]) for r in range(4))) for c in range(3)])
sm0 = int(smid[0, 0]); ctas = [c for c in range(nseg) if smid[c, 0] == sm0]
print("CTAs on SM", sm0, ":", ctas, "warpids:", [[int(wid[c, r]) for r in range(4)] for c in ctas])
