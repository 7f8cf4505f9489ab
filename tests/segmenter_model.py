"""numpy restatement of the CW-gap segmenter (include/rfid_b200.h, csrc/rx_ingest.cuh) -- test infrastructure.

The reference has no segmenter; its role in the tests is (1) to pin the CUDA segmenter's
table sample for sample and (2) to let the CPU suite check the segmentation rule against the golden capture.
"""
import numpy as np

LEVEL_HEAD = 1 << 21


def bursts(iq, adc_rate=2_000_000, level_frac=0.5, gap_us=400.0):
    x = np.ascontiguousarray(iq, dtype=np.complex64)
    re, im = x.real.astype(np.float32), x.imag.astype(np.float32)
    p = (re * re + im * im).astype(np.float32)
    n0 = min(x.size, LEVEL_HEAD)
    level = np.float32(np.sqrt(p[:n0]).astype(np.float64).sum() / n0) if n0 else np.float32(0)
    thr = np.float32(level_frac) * level
    thr2 = np.float32(thr * thr)
    low = p < thr2
    prev = np.concatenate(([False], low[:-1]))
    falls = np.flatnonzero(low & ~prev)
    lows = np.flatnonzero(low)
    gap = int(float(np.float32(gap_us)) * 1e-6 * adc_rate)
    # last low sample strictly before each falling edge
    k = np.searchsorted(lows, falls, side="left") - 1
    prev_low = np.where(k >= 0, lows[np.maximum(k, 0)], -1)
    start = (prev_low < 0) | (falls - prev_low - 1 >= gap)
    pos = falls[start]
    rank = np.flatnonzero(start)
    pulses = np.diff(np.concatenate((rank, [falls.size])))
    return pos, pulses


def segment_table(iq, adc_rate=2_000_000, decim=5, level_frac=0.5, gap_us=400.0, lead_us=300.0, min_pulses=6, commands=2):
    n = int(np.asarray(iq).size)
    pos, pulses = bursts(iq, adc_rate, level_frac, gap_us)
    cmd = pos[pulses >= min_pulses]
    lead = int(float(np.float32(lead_us)) * 1e-6 * adc_rate)
    ns = (len(cmd) + commands - 1) // commands if len(cmd) else (1 if n else 0)
    out = []
    for j in range(ns):
        start = 0
        if j > 0:
            start = max(int(cmd[j * commands]) - lead, 0)
            start -= start % decim
        end = int(cmd[(j + 1) * commands]) if (j + 1) * commands < len(cmd) else n
        out.append((start, end - start))
    return out, cmd
