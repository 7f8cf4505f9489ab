"""Synthetic Gen2 RX captures (seeded, torch, device-agnostic).

Signal model (calibrated on the reference's recorded capture
gr-rfid/misc/data/file_source_test):

    rx[n] = (L + sum_k g_k * b_k[n]) * tx[n] + w[n]            (raw rate, complex64)

* tx: the reader's PIE envelope exactly as reader_impl builds it at 1 us resolution
  (reader_impl.cc:51-125,251-344: delimiter 12 us, data0 12+12, data1 36+12, RTcal 60+12,
  TRcal 188+12, Query = preamble + 22 bits incl. CRC-5, QueryRep = frame-sync + 0000,
  ACK = frame-sync + 01 + RN16), band-limited by a short FIR measured on the recording;
* L: carrier leakage 0.2846-0.0349j; g_k: per-tag backscatter gain (|g| ~ 0.0227);
* b_k: FM0 half-symbol levels at BLF 40 kHz: TAG_PREAMBLE (global_vars.h:136), data, dummy '1'
  (global_vars.h:104-107); reply starts T1 ~ 248 us after the command's last rising edge;
  the tag's own symbol clock is off by up to +-0.8 %;
* w: complex white Gaussian noise, sigma ~ 0.003 per component.

One *segment* = one inventory slot: lead-in CW, Query (slot 0) or QueryRep, RN16 reply, ACK,
EPC reply, tail CW.  With n_tags > 1 several tags may answer in the same slot (collision): all of
them send an RN16, the reader ACKs the strongest one and only that tag sends its EPC.

This module is product code (the benchmark workload generator); it does not touch oracle/.
"""
import math

import numpy as np
import torch

TAG_PREAMBLE = (1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1)   # include/rfid/global_vars.h:136
LEAK = complex(0.2846, -0.0349)
TAG_GAIN = 0.0227
TAG_DIR = math.atan2(0.192, -0.981)
NOISE_SIGMA = 0.0030
# step response of the TX/RX chain measured on file_source_test (falling edge, raw 2 MS/s samples)
EDGE_FIR_2MSPS = (-0.07, 0.32, 0.52, 0.155, 0.005, 0.03, 0.01, 0.01, 0.01, 0.005, 0.005)
FLOOR = 0.004

# durations in us: include/rfid/global_vars.h:90-97, reader_impl.cc:51-71
PW_US, DELIM_US, DATA0_US, DATA1_US, RTCAL_US, TRCAL_US = 12, 12, 24, 48, 72, 200
T1_TAG_US = 248.0
CW_QUERY_US = 240 + 480 + 575        # n_cwquery_s, reader_impl.cc:69
CW_ACK_US = 3 * 240 + 480 + 3375     # n_cwack_s,   reader_impl.cc:70
HALF_SYMBOL_US = 12.5                # BLF 40 kHz


def crc16_gen2(data: bytes) -> int:
    """CRC-16/CCITT as the tag computes it (EPC Gen2 annex F; same algorithm the
    reference checks in tag_decoder_impl.cc:424-440)."""
    crc = 0xFFFF
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return (~crc) & 0xFFFF


def crc5_gen2(bits17):
    """CRC-5 of a 17-bit Query (x^5+x^3+1, preset 01001), MSB first (reader_impl.cc:383-443)."""
    reg = 0b01001
    for b in bits17:
        fb = ((reg >> 4) & 1) ^ int(b)
        reg = (reg << 1) & 0x1F
        if fb:
            reg ^= 0b01001
    return [(reg >> k) & 1 for k in (4, 3, 2, 1, 0)]


def query_bits(fixed_q: int):
    head = [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0] + [(fixed_q >> (3 - b)) & 1 for b in range(4)]
    return head + crc5_gen2(head)


def make_epc_frames(epc96: np.ndarray) -> np.ndarray:
    """[n,12] uint8 EPC words -> [n,16] uint8 frames PC(0x3000) + EPC + CRC-16."""
    n = epc96.shape[0]
    out = np.zeros((n, 16), dtype=np.uint8)
    out[:, 0] = 0x30
    out[:, 2:14] = epc96
    for i in range(n):
        c = crc16_gen2(bytes(out[i, :14]))
        out[i, 14] = c >> 8
        out[i, 15] = c & 0xFF
    return out


def _pulse_starts(kind_is_query: torch.Tensor, qbits, ack_bits: torch.Tensor, t_cmd: float):
    """Start times (us) of every low pulse of the slot's two commands.

    Returns (starts [S,P] float64 with +inf padding, cmd1_end [S], cmd2_end [S], ack_start[S])."""
    S = ack_bits.shape[0]
    dev = ack_bits.device
    # command 1: Query (26 pulses) or QueryRep (7 pulses); low pulse = last PW of each symbol,
    # the delimiter is a low pulse of its own
    q_sym = [DATA0_US, RTCAL_US, TRCAL_US] + [DATA1_US if b else DATA0_US for b in qbits]
    q_start = [0.0]
    t = float(DELIM_US)
    for d in q_sym:
        t += d
        q_start.append(t - PW_US)
    q_end = t
    r_sym = [DATA0_US, RTCAL_US] + [DATA0_US] * 4
    r_start = [0.0]
    t = float(DELIM_US)
    for d in r_sym:
        t += d
        r_start.append(t - PW_US)
    r_end = t
    P1 = len(q_start)
    inf = float("inf")
    q_t = torch.tensor(q_start, dtype=torch.float64, device=dev)
    r_t = torch.tensor(r_start + [inf] * (P1 - len(r_start)), dtype=torch.float64, device=dev)
    c1 = torch.where(kind_is_query[:, None], q_t[None, :], r_t[None, :]) + t_cmd
    cmd1_end = torch.where(kind_is_query, torch.tensor(q_end, dtype=torch.float64, device=dev),
                           torch.tensor(r_end, dtype=torch.float64, device=dev)) + t_cmd
    # command 2: ACK = frame-sync + 18 bits, CW_QUERY_US after command 1 (reader_impl.cc:283-285,290-320)
    ack_start = cmd1_end + CW_QUERY_US
    dur = torch.where(ack_bits > 0, float(DATA1_US), float(DATA0_US)).to(torch.float64)   # [S,18]
    ends = DELIM_US + DATA0_US + RTCAL_US + torch.cumsum(dur, dim=1)                      # end of each bit symbol
    a_start = torch.cat([torch.zeros(S, 1, dtype=torch.float64, device=dev),
                         torch.full((S, 1), DELIM_US + DATA0_US - PW_US, dtype=torch.float64, device=dev),
                         torch.full((S, 1), DELIM_US + DATA0_US + RTCAL_US - PW_US, dtype=torch.float64, device=dev),
                         ends - PW_US], dim=1)
    c2 = a_start + ack_start[:, None]
    cmd2_end = ack_start + ends[:, -1]
    return torch.cat([c1, c2], dim=1), cmd1_end, cmd2_end, ack_start


def _fm0_levels(bits: torch.Tensor) -> torch.Tensor:
    """bits [S,B] (0/1) -> half-symbol levels [S, 12 + 2*(B+1)] : preamble, data, dummy 1."""
    S, B = bits.shape
    dev = bits.device
    pre = torch.tensor(TAG_PREAMBLE, dtype=torch.int8, device=dev)[None, :].expand(S, 12)
    allbits = torch.cat([bits.to(torch.int8), torch.ones(S, 1, dtype=torch.int8, device=dev)], dim=1)   # + dummy
    # first half of symbol j = !(last level before it); second half = first (bit 1) or !first (bit 0).
    # last level flips once per bit-1 symbol and twice (= not at all) per bit-0 symbol.
    flips = torch.cumsum(allbits.to(torch.int32), dim=1)              # number of '1' symbols up to and incl. j
    last_prev = (1 + flips - allbits.to(torch.int32)) & 1             # level at the end of symbol j-1 (preamble ends high)
    first = 1 - last_prev
    second = torch.where(allbits > 0, first, 1 - first)
    data = torch.stack([first, second], dim=2).reshape(S, 2 * (B + 1)).to(torch.int8)
    return torch.cat([pre, data], dim=1)


def make_capture(n_segments: int, *, adc_rate: int = 2_000_000, segment_us: float = 8480.0, lead_us: float = 400.0,
                 fixed_q: int = 0, n_tags: int = 1, seed: int = 1234, first_segment: int = 0,
                 noise_sigma: float = NOISE_SIGMA, tag_gain: float = TAG_GAIN, clock_ppm_pct: float = 0.8,
                 device="cpu", chunk_segments: int = 4096, out: torch.Tensor = None):
    """Generate `n_segments` inventory slots back to back.

    Returns dict(iq=complex64 tensor [n_segments*seg_len], segments=np structured array (abi.SEGMENT_DTYPE),
    truth=dict(rn16 [S] int, epc [S,16] uint8, n_replies [S], is_query [S])).
    Segment ids (first_segment + i) seed the per-segment randomness, so a rank that generates only its
    shard produces exactly the samples the single-process run would.
    """
    from . import abi
    sps = adc_rate / 1e6
    seg_len = int(round(segment_us * sps))
    dev = torch.device(device)
    S = n_segments
    if out is None:
        out = torch.empty(S * seg_len, dtype=torch.complex64, device=dev)
    iq = out.view(S, seg_len)
    slots = 1 << fixed_q
    qb = query_bits(fixed_q)

    # ---------- per-segment random draws (CPU generator keyed by segment id => shard-invariant) ----------
    ids = np.arange(first_segment, first_segment + S, dtype=np.int64)
    rng_all = np.random.Generator(np.random.Philox(key=seed))
    # draw per-round tag -> slot assignments deterministically from (seed, round)
    rounds = ids // slots
    slot_in_round = ids % slots
    K = max(1, n_tags)
    rn16 = np.zeros((S, K), dtype=np.int64)
    present = np.zeros((S, K), dtype=bool)
    gain = np.zeros((S, K), dtype=np.complex128)
    clk = np.zeros((S, K), dtype=np.float64)
    jit = np.zeros((S, K), dtype=np.float64)
    for i in range(S):
        r = np.random.Generator(np.random.Philox(key=[seed, int(rounds[i])]))
        tag_slots = r.integers(0, slots, size=K)
        tag_amp = tag_gain * r.uniform(0.6, 1.4, size=K)
        tag_ph = TAG_DIR + r.uniform(-0.6, 0.6, size=K)
        s = np.random.Generator(np.random.Philox(key=[seed + 1, int(ids[i])]))
        rn16[i] = s.integers(0, 65536, size=K)
        clk[i] = s.uniform(-clock_ppm_pct, clock_ppm_pct, size=K) / 100.0
        jit[i] = s.uniform(0.0, 2.0, size=K)
        if n_tags > 0:
            present[i] = tag_slots == slot_in_round[i]
        gain[i] = tag_amp * np.exp(1j * tag_ph)
    del rng_all
    amp = np.where(present, np.abs(gain), 0.0)
    strongest = amp.argmax(axis=1)
    n_replies = present.sum(axis=1)
    ack_rn16 = np.where(n_replies > 0, rn16[np.arange(S), strongest], 0)
    # tag EPC: 96-bit word derived from (seed, tag index); last byte = tag index + 0x27 (the recording's tag is 0x27)
    epc_tab = np.zeros((K, 12), dtype=np.uint8)
    for k in range(K):
        t = np.random.Generator(np.random.Philox(key=[seed + 2, k]))
        epc_tab[k] = t.integers(0, 256, size=12)
        epc_tab[k, 11] = (0x27 + k) & 0xFF
    frames = make_epc_frames(epc_tab)                                       # [K,16]
    epc_frame = frames[strongest]                                           # [S,16]
    ack_bits_np = np.concatenate([np.tile(np.array([[0, 1]], dtype=np.int8), (S, 1)),
                                  ((ack_rn16[:, None] >> np.arange(15, -1, -1)[None, :]) & 1).astype(np.int8)], axis=1)

    hs = HALF_SYMBOL_US * sps                                               # raw samples per half symbol
    fir = torch.tensor(EDGE_FIR_2MSPS, dtype=torch.float32, device=dev)
    if abs(sps - 2.0) > 1e-9:   # stretch the edge filter to the same duration at other rates
        L = max(3, int(round(len(EDGE_FIR_2MSPS) * sps / 2.0)))
        fir = torch.nn.functional.interpolate(fir[None, None, :], size=L, mode="linear", align_corners=True)[0, 0]
        fir = fir / fir.sum()
    n_idx = torch.arange(seg_len, device=dev, dtype=torch.float32)[None, :]

    for c0 in range(0, S, chunk_segments):
        c1 = min(S, c0 + chunk_segments)
        n = c1 - c0
        is_q = torch.from_numpy(slot_in_round[c0:c1] == 0).to(dev)
        ackb = torch.from_numpy(ack_bits_np[c0:c1]).to(dev)
        starts, cmd1_end, cmd2_end, _ = _pulse_starts(is_q, qb, ackb, lead_us)
        # ---- TX envelope: +1/-1 deltas at pulse starts/ends, cumsum, FIR ----
        P = starts.shape[1]
        finite = torch.isfinite(starts)
        s_idx = torch.where(finite, torch.round(starts * sps), torch.zeros_like(starts)).long().clamp(0, seg_len)
        e_idx = torch.where(finite, torch.round((starts + PW_US) * sps), torch.zeros_like(starts)).long().clamp(0, seg_len)
        delta = torch.zeros(n, seg_len + 1, dtype=torch.float32, device=dev)
        w = finite.to(torch.float32)
        delta.scatter_add_(1, s_idx, w)
        delta.scatter_add_(1, e_idx, -w)
        low = torch.cumsum(delta[:, :seg_len], dim=1)
        env = 1.0 - (1.0 - FLOOR) * low.clamp(0, 1)
        pad = fir.numel() - 1
        env = torch.nn.functional.conv1d(torch.nn.functional.pad(env[:, None, :], (pad - 1, 1), value=1.0),
                                         fir.flip(0)[None, None, :])[:, 0, :]
        # ---- tag backscatter ----
        refl = torch.zeros(n, seg_len, dtype=torch.complex64, device=dev)
        for k in range(K):
            pres = torch.from_numpy(present[c0:c1, k]).to(dev)
            if not bool(pres.any()):
                continue
            g = torch.from_numpy(gain[c0:c1, k].astype(np.complex64)).to(dev)
            hk = torch.from_numpy((hs * (1.0 + clk[c0:c1, k])).astype(np.float32)).to(dev)[:, None]
            j = torch.from_numpy(jit[c0:c1, k]).to(dev)
            # RN16 reply (every present tag)
            rbits = torch.from_numpy(((rn16[c0:c1, k][:, None] >> np.arange(15, -1, -1)[None, :]) & 1).astype(np.int8)).to(dev)
            lv = _fm0_levels(rbits)
            t0 = ((cmd1_end + T1_TAG_US + j) * sps).to(torch.float32)[:, None]
            h = torch.floor((n_idx - t0) / hk).long()
            ok = (h >= 0) & (h < lv.shape[1]) & pres[:, None]
            b = torch.gather(lv, 1, h.clamp(0, lv.shape[1] - 1)).to(torch.float32) * ok
            refl += g[:, None] * b
            # EPC reply (only the tag the reader ACKed)
            sel = pres & torch.from_numpy(strongest[c0:c1] == k).to(dev)
            if bool(sel.any()):
                fb = np.unpackbits(frames[k])[None, :].repeat(n, 0).astype(np.int8)
                lv = _fm0_levels(torch.from_numpy(fb).to(dev))
                t0 = ((cmd2_end + T1_TAG_US + j) * sps).to(torch.float32)[:, None]
                h = torch.floor((n_idx - t0) / hk).long()
                ok = (h >= 0) & (h < lv.shape[1]) & sel[:, None]
                b = torch.gather(lv, 1, h.clamp(0, lv.shape[1] - 1)).to(torch.float32) * ok
                refl += g[:, None] * b
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed) * 1_000_003 + int(ids[c0]))
        noise = torch.randn(n, seg_len, 2, generator=gen, device=dev, dtype=torch.float32) * noise_sigma
        iq[c0:c1] = (LEAK + refl) * env + torch.view_as_complex(noise)

    segs = abi.make_segments(np.arange(S, dtype=np.uint64) * seg_len, np.full(S, seg_len, dtype=np.uint32))
    truth = {"rn16": ack_rn16.astype(np.int64), "epc": epc_frame, "n_replies": n_replies,
             "is_query": slot_in_round == 0, "segment_len": seg_len}
    return {"iq": out, "segments": segs, "truth": truth}
