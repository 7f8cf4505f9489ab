#!/usr/bin/env python
"""Regenerates tests/golden/* from /root/reference (run in the build container only).

  file_source_test.c64.xz   the reference's recorded RX capture (gr-rfid/misc/data/file_source_test,
                            1,247,958 complex64 @ 2 MS/s), xz-compressed verbatim: the cfg1 parity input
  file_sink_commands.json   Query / ACK bit strings decoded from the reference author's committed TX output
                            gr-rfid/misc/data/file_sink (PIE: data0 = 24 samples fall-to-fall, data1 = 48;
                            reader_impl.cc:51-71,84-125) -- 72 Queries and 71 ACKs whose payloads are the
                            RN16s the reference decoded on that run
  readme_expected.txt       the known-answer block of README.md:46-53
  cfg1_ref_records.npy      rfid_b200_window_result records produced by oracle/_ref (the reference's own
                            blocks) on file_source_test; cfg1_ref_stats.json the READER_STATS + print_results text
  cfg1_q4_*.                the same with FIXED_Q = 4
"""
import json
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def decode_pie(tx):
    """TX envelope (0/1 floats @1 MS/s) -> list of (kind, bitstring)."""
    lvl = tx > 0.5
    fall = np.nonzero(lvl[:-1] & ~lvl[1:])[0] + 1   # first low sample of each pulse
    # group falls into commands: gaps > 400 samples separate commands
    cmds, cur = [], [fall[0]]
    for a, b in zip(fall[:-1], fall[1:]):
        if b - a > 400:
            cmds.append(cur)
            cur = []
        cur.append(b)
    cmds.append(cur)
    out = []
    for c in cmds:
        d = np.diff(c)
        iv = list(d)
        # fall-to-fall intervals: delimiter->data0 low = 24, RTcal = 72, TRcal = 200 (Query only),
        # then one interval per bit (every PIE symbol ends with its low pulse): 24 = '0', 48 = '1'
        assert iv[0] == 24 and iv[1] == 72, iv[:4]
        k = 2
        kind = "framesync"
        if iv[k] == 200:
            kind = "preamble"
            k += 1
        assert all(x in (24, 48) for x in iv[k:]), iv
        bits = "".join("1" if x == 48 else "0" for x in iv[k:])
        out.append((kind, bits))
    return out


def main():
    src = os.path.join(REF, "gr-rfid/misc/data/file_source_test")
    raw = open(src, "rb").read()
    with open(os.path.join(HERE, "file_source_test.c64.xz"), "wb") as f:
        f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))

    sink = np.fromfile(os.path.join(REF, "gr-rfid/misc/data/file_sink"), dtype=np.complex64).real
    amp = sink.max()
    cmds = decode_pie(sink / amp)
    queries = [b for k, b in cmds if k == "preamble"]
    acks = [b for k, b in cmds if k == "framesync"]
    rn16 = ["%04x" % int(b[2:], 2) for b in acks]
    json.dump({"amplitude": float(amp), "n_lead_cw": int(np.argmax(sink / amp < 0.5)),
               "queries": queries, "acks": acks, "rn16": rn16},
              open(os.path.join(HERE, "file_sink_commands.json"), "w"), indent=1)

    readme = open(os.path.join(REF, "README.md")).read().splitlines()
    open(os.path.join(HERE, "readme_expected.txt"), "w").write("\n".join(readme[45:53]) + "\n")

    from oracle.refflow import RefFlow
    iq = np.frombuffer(raw, dtype=np.complex64)
    for q, tag in ((0, "cfg1"), (4, "cfg1_q4")):
        r = RefFlow(q).run_stream(iq)
        np.save(os.path.join(HERE, tag + "_ref_records.npy"), r["records"])
        s = r["stats"]
        json.dump({"text": r["text"], "n_queries_sent": s.n_queries_sent,
                   "cur_inventory_round": s.cur_inventory_round, "cur_slot_number": s.cur_slot_number,
                   "n_epc_correct": s.n_epc_correct, "tag_reads": {str(k): v for k, v in s.tag_map().items()},
                   "n_windows": r["n_windows"]},
                  open(os.path.join(HERE, tag + "_ref_stats.json"), "w"), indent=1)
    print("golden regenerated:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
