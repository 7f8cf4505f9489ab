"""ctypes mirror of include/rfid_b200.h (structures and constants only).

Kept separate from the library loader so that test infrastructure can
share the record layout without importing the CUDA library.
"""
import ctypes as C

import numpy as np

RN16 = 0
EPC = 1
MAX_TAGS = 256

OK, EINVAL, ENODEV, ENOMEM, ECUDA, ECAPACITY = 0, -1, -2, -3, -4, -5


class Params(C.Structure):
    """rfid_b200_params (defaults: apps/reader.py:52-65, include/rfid/global_vars.h:72-143)."""
    _fields_ = [("adc_rate", C.c_int32), ("decim", C.c_int32), ("ntaps", C.c_int32),
                ("fixed_q", C.c_int32), ("max_queries", C.c_int32), ("max_tags", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32)]


class Segment(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint32), ("reserved", C.c_uint32)]


class WindowResult(C.Structure):
    _fields_ = [("segment", C.c_int32), ("window", C.c_int32), ("open_index", C.c_int32),
                ("length", C.c_int32), ("kind", C.c_int32), ("sync_index", C.c_int32),
                ("score", C.c_float), ("h_re", C.c_float), ("h_im", C.c_float), ("T", C.c_float),
                ("crc_ok", C.c_int32), ("tag_id", C.c_int32), ("bits", C.c_uint8 * 16)]


class Stats(C.Structure):
    _fields_ = [("n_queries_sent", C.c_int32), ("cur_inventory_round", C.c_int32),
                ("cur_slot_number", C.c_int32), ("max_slot_number", C.c_int32),
                ("n_epc_correct", C.c_int32), ("n_windows", C.c_int32), ("terminated", C.c_int32),
                ("n_unique_tags", C.c_int32), ("tag_id", C.c_int32 * MAX_TAGS),
                ("tag_reads", C.c_int32 * MAX_TAGS)]

    def tag_map(self):
        n = min(self.n_unique_tags, MAX_TAGS)
        return {int(self.tag_id[i]): int(self.tag_reads[i]) for i in range(n)}


class Segmenter(C.Structure):
    """rfid_b200_segmenter (include/rfid_b200.h): CW-gap segmenter settings."""
    _fields_ = [("level_frac", C.c_float), ("gap_us", C.c_float), ("lead_us", C.c_float),
                ("min_pulses", C.c_int32), ("commands_per_segment", C.c_int32), ("reserved", C.c_int32 * 3)]


class TxCommand(C.Structure):
    _fields_ = [("kind", C.c_int32), ("arg", C.c_int32)]


class SimParams(C.Structure):
    """rfid_b200_sim_params (include/rfid_b200.h): inventory-slot simulator settings."""
    _fields_ = [("seed", C.c_uint64), ("n_tags", C.c_int32), ("closed_loop", C.c_int32), ("dac_rate", C.c_int32),
                ("segment_us", C.c_float), ("lead_us", C.c_float), ("noise_sigma", C.c_float), ("tag_gain", C.c_float),
                ("tag_phase", C.c_float), ("clock_pct", C.c_float), ("leak_re", C.c_float), ("leak_im", C.c_float),
                ("floor_level", C.c_float), ("reserved", C.c_int32 * 2)]


TX_START, TX_QUERY, TX_QUERY_REP, TX_ACK, TX_CW, TX_NAK, TX_POWER_DOWN, TX_QUERY_ADJUST = range(8)
TX_COMMAND_DTYPE = np.dtype([("kind", "<i4"), ("arg", "<i4")])
SIM_TRUTH_DTYPE = np.dtype([("is_query", "<i4"), ("n_replies", "<i4"), ("strongest_rn16", "<i4"), ("acked_rn16", "<i4"),
                            ("replier", "<i4"), ("reserved", "<i4", (3,)), ("epc", "u1", (16,))])
assert C.sizeof(SimParams) == 64 and SIM_TRUTH_DTYPE.itemsize == 48 and C.sizeof(TxCommand) == 8
assert C.sizeof(Segmenter) == 32
assert C.sizeof(WindowResult) == 64
assert C.sizeof(Segment) == 16
assert C.sizeof(Params) == 32

#: numpy view of rfid_b200_window_result (same 64-byte layout)
RESULT_DTYPE = np.dtype([("segment", "<i4"), ("window", "<i4"), ("open_index", "<i4"), ("length", "<i4"),
                         ("kind", "<i4"), ("sync_index", "<i4"), ("score", "<f4"), ("h_re", "<f4"),
                         ("h_im", "<f4"), ("T", "<f4"), ("crc_ok", "<i4"), ("tag_id", "<i4"),
                         ("bits", "u1", (16,))])
SEGMENT_DTYPE = np.dtype([("offset", "<u8"), ("length", "<u4"), ("reserved", "<u4")])
assert RESULT_DTYPE.itemsize == 64 and SEGMENT_DTYPE.itemsize == 16


def make_segments(offsets, lengths):
    segs = np.zeros(len(offsets), dtype=SEGMENT_DTYPE)
    segs["offset"] = np.asarray(offsets, dtype=np.uint64)
    segs["length"] = np.asarray(lengths, dtype=np.uint32)
    return segs


def bits_hex(rec):
    """hex string of a record's decoded bits (4 hex digits for an RN16, 32 for an EPC)."""
    b = bytes(bytearray(rec["bits"]))
    return b[:2].hex() if int(rec["kind"]) == RN16 else b.hex()
