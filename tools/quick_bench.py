#!/usr/bin/env python
"""Developer aid: parity spot-check + kernel timing at a few batch sizes (GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import abi, capi, synth  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rx = capi.Gen2Rx()
    cap = synth.make_capture(160, seed=3, device=dev)
    segs = capi.segments_to_device(cap["segments"], dev)
    res, cnt = rx.decode_capture(cap["iq"], segs, 4)
    torch.cuda.synchronize()
    recs, counts = capi.results_to_numpy(res, cnt, 4)
    orecs, ocounts, _ = Oracle().decode_segments(cap["iq"].cpu().numpy(), cap["segments"], max_per_seg=4)
    print("parity 160 segs:", "BIT-EXACT" if recs.tobytes() == orecs.tobytes() and (counts == ocounts).all() else "MISMATCH")
    sizes = [int(x) for x in (sys.argv[1:] or ["1000", "4000", "16000"])]
    for nseg in sizes:
        caps = [synth.make_capture(nseg, seed=10 + b, device=dev)["iq"] for b in range(2 if nseg > 4000 else 4)]
        segs = capi.segments_to_device(abi.make_segments(np.arange(nseg, dtype=np.uint64) * 16960, [16960] * nseg), dev)
        res = torch.zeros((nseg * 2, 64), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(nseg, dtype=torch.int32, device=dev)
        for i in range(3):
            rx.decode_capture(caps[i % len(caps)], segs, 2, res, cnt)
        torch.cuda.synchronize()
        rx.enable_kernel_timing(True)
        rx.kernel_time(True)
        it = 20 if nseg <= 4000 else 6
        for i in range(it):
            rx.decode_capture(caps[i % len(caps)], segs, 2, res, cnt)
        torch.cuda.synchronize()
        ms, n = rx.kernel_time(True)
        rx.enable_kernel_timing(False)
        us = 1e3 * ms / n
        gbs = nseg * 16960 * 8 / (us * 1e-6) / 1e9
        ok = int((capi.results_to_numpy(res, cnt, 2)[0][:, 1]["crc_ok"] == 1).sum())
        print("nseg %6d: kernel %8.1f us  %7.1f GB/s  %5.1f%% of 6571.6  (%.1f GS/s)  epc_ok %d" %
              (nseg, us, gbs, 100 * gbs / 6571.6, gbs / 8, ok))
        del caps


if __name__ == "__main__":
    main()
