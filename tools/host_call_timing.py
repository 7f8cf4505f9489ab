#!/usr/bin/env python
"""Developer aid: wall-clock time of successive rfid_b200_decode_capture_host calls from pinned memory (GPU box)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import abi, capi, synth  # noqa: E402
dev = torch.device("cuda:0")
rx = capi.Gen2Rx()
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cap = synth.make_capture(nseg, seed=5, device=dev)
h_iq = torch.empty(cap["iq"].numel(), dtype=torch.complex64).pin_memory(); h_iq.copy_(cap["iq"])
segs_np = cap["segments"]
h_segs = torch.from_numpy(np.ascontiguousarray(segs_np).view(np.uint8).copy()).pin_memory()
h_res = torch.zeros((nseg * 2, 64), dtype=torch.uint8).pin_memory()
h_cnt = torch.zeros(nseg, dtype=torch.int32).pin_memory()
ts = []
for i in range(12):
    t0 = time.perf_counter()
    rx.decode_capture_host_ptr(h_iq.data_ptr(), h_iq.numel(), h_segs.data_ptr(), nseg, 2, h_res.data_ptr(), h_cnt.data_ptr())
    ts.append((time.perf_counter() - t0) * 1e3)
print("nseg", nseg, "ms per call:", " ".join("%.2f" % t for t in ts), "launches", rx.last_launches() if hasattr(rx, "last_launches") else "?")
recs = h_res.numpy().reshape(-1).view(abi.RESULT_DTYPE).reshape(nseg, 2)
print("epc ok", int((recs[:, 1]["crc_ok"] == 1).sum()))
