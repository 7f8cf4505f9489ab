#!/usr/bin/env python
"""Developer aid: smallest-first parity runs of the capture kernel against the oracle (GPU box); run under `timeout`."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import abi, capi, synth  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

dev = torch.device("cuda:0")
rx = capi.Gen2Rx()
for nseg in [int(x) for x in (sys.argv[1:] or ["1", "2", "7", "160"])]:
    cap = synth.make_capture(nseg, seed=3, device=dev)
    segs = capi.segments_to_device(cap["segments"], dev)
    print("nseg", nseg, "launch", flush=True)
    res, cnt = rx.decode_capture(cap["iq"], segs, 4)
    torch.cuda.synchronize()
    recs, counts = capi.results_to_numpy(res, cnt, 4)
    orecs, ocounts, _ = Oracle().decode_segments(cap["iq"].cpu().numpy(), cap["segments"], max_per_seg=4)
    ok = recs.tobytes() == orecs.tobytes() and (counts == ocounts).all()
    print("nseg", nseg, "BIT-EXACT" if ok else "MISMATCH", "counts", counts[:8], ocounts[:8], flush=True)
    if not ok:
        bad = [i for i in range(nseg) if recs[i].tobytes() != orecs[i].tobytes()]
        print(" first bad segment", bad[:5])
        i = bad[0] if bad else 0
        for f in recs.dtype.names:
            if recs[i][f].tobytes() != orecs[i][f].tobytes(): print("  field", f, recs[i][f], orecs[i][f])
        break
