#!/usr/bin/env python
"""Developer aid (GPU box): per-call latency of the block-mode entry points the GNU Radio host blocks use -- the
prerequisite number for real-time operation (the reader must answer within T2 <= 500 us, include/rfid/global_vars.h:93).
Writes one JSON object (profiles/r2_block_latency.json when run through gpurun_out)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gen2_uhf_rfid_reader_b200 import capi, synth
from oracle.pyoracle import Oracle

rx = capi.Gen2Rx()
cap = synth.make_capture(8, seed=11, device="cpu")
y = Oracle().mf(cap["iq"].numpy())                    # the flowgraph's own matched filter runs upstream of the gate
raw = cap["iq"].numpy()


def timed(fn, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[len(ts) // 10:]) * 1e6
    return {"median_us": float(np.median(ts)), "p90_us": float(np.percentile(ts, 90)), "min_us": float(ts.min()), "calls": int(ts.size)}


out = {"what": "host pointer in -> kernel -> host pointer out, one call, one stream synchronisation", "chunk_samples": 4096}
pos = [0]
def gate_call():
    c = y[pos[0]: pos[0] + 4096]
    if c.size < 4096:
        pos[0] = 0; c = y[:4096]
    r = rx.gate_work(c, seek=0)
    pos[0] += max(1, r["consumed"])
out["gate_work_4096"] = timed(gate_call, 300)
win_rn = np.ascontiguousarray(y[1000:1000 + rx.len_rn16]); win_epc = np.ascontiguousarray(y[2000:2000 + rx.len_epc])
out["decoder_work_rn16"] = timed(lambda: rx.decoder_work(0, win_rn), 200)
out["decoder_work_epc"] = timed(lambda: rx.decoder_work(1, win_epc), 200)
rpos = [0]
def mf_call():
    c = raw[rpos[0]: rpos[0] + 20480]
    if c.size < 20480:
        rpos[0] = 0; c = raw[:20480]
    rx.mf_work(c); rpos[0] += 20480
out["mf_work_20480_raw"] = timed(mf_call, 200)
out["t2_budget_us"] = 500
print(json.dumps(out))
