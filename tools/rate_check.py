#!/usr/bin/env python
"""Developer aid: parity at other sample rates / tap counts (exercises the generic kernel paths)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import abi, capi, synth
from oracle.pyoracle import Oracle

ok_all = True
for adc, ntaps in ((2000000, 25), (1000000, 13), (1000000, 12), (2000000, 20), (4000000, 50), (6000000, 75), (8000000, 100)):
    rx = capi.Gen2Rx(adc_rate=adc, ntaps=ntaps)
    O = Oracle(adc_rate=adc, ntaps=ntaps)
    cap = synth.make_capture(24, seed=4, adc_rate=adc)
    iq = cap["iq"].numpy()
    recs, counts = rx.decode_capture_host(iq, cap["segments"], max_windows=4)
    orecs, ocounts, _ = O.decode_segments(iq, cap["segments"], max_per_seg=4)
    same = recs.tobytes() == orecs.tobytes() and (counts == ocounts).all()
    good = int((orecs[:, 1]["crc_ok"] == 1).sum())
    print("adc %d ntaps %3d: windows %3d  epc_ok(ref) %2d/24  %s" % (adc, ntaps, counts.sum(), good, "BIT-EXACT" if same else "MISMATCH"))
    if not same:
        for f in recs.dtype.names:
            if recs[f].tobytes() != orecs[f].tobytes():
                bad = np.nonzero((recs[f] != orecs[f]).reshape(recs.size, -1).any(axis=1))[0]
                print("   field", f, "differs in", bad.size, "first", bad[:3], recs.reshape(-1)[f][bad[:2]], orecs.reshape(-1)[f][bad[:2]])
    ok_all &= same
print("ALL OK" if ok_all else "FAILURES")
