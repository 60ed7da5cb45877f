"""The wire-decode oracle (oracle/wire_oracle.c) against the committed golden vectors
(tests/golden/wire/messages.json: python-protobuf encodings + verdicts of eraftpb.Message,
scripts/gen_wire_golden.py).  The reference has no wire goldens and its codecs (rust-protobuf 2 /
prost 0.7) are third-party: "parity unpinned by the reference", pinned to the published format."""
import json
import os

import numpy as np

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def vectors():
    with open(os.path.join(HERE, "golden", "wire", "messages.json")) as f:
        return json.load(f)["vectors"]


def test_every_golden_message_decodes_to_the_same_fields():
    n_ok = n_bad = 0
    for v in vectors():
        got = O.wire_decode_message(bytes.fromhex(v["hex"]))
        if not v["ok"]:
            assert got is None, (v["kind"], v["hex"])
            n_bad += 1
            continue
        assert got is not None, (v["kind"], v["hex"])
        want = dict(v["fields"])
        want["msg_type"] &= 0xFFFFFFFF
        assert got == want, (v["kind"], v["hex"], got, want)
        n_ok += 1
    assert n_ok > 350 and n_bad >= 15


def test_batch_statuses_and_records():
    """wo_decode_batch over frames built from the golden messages: type / term / needs-log / malformed / dup
    statuses and the record fields of the accepted AppendResponses."""
    vs = vectors()
    frames, want = [], []
    seen = set()
    n_groups = 512
    terms = np.zeros(n_groups, dtype=np.uint64)
    terms[::2] = 7     # even groups check the term, odd groups leave it to the caller (0)
    for i, v in enumerate(vs):
        g, slot = (i * 7) % n_groups, i % 8
        if i % 11 == 0 and frames:      # a second message for an earlier cell
            g, slot = want[-1][1], want[-1][2]
        hdr = np.array([g << 4 | slot], dtype="<u4").tobytes()
        frames.append(hdr + bytes.fromhex(v["hex"]))
        if not v["ok"]:
            st = O.WIRE_MALFORMED
        else:
            f = v["fields"]
            if (f["msg_type"] & 0xFFFFFFFF) != 4:
                st = O.WIRE_SKIP_TYPE
            elif terms[g] != 0 and f["term"] != terms[g]:
                st = O.WIRE_TERM
            elif f["reject"] and f["log_term"] > 0:
                st = O.WIRE_NEEDS_LOG
            elif (g, slot) in seen:
                st = O.WIRE_DUP
            else:
                st = O.WIRE_OK
                seen.add((g, slot))
        want.append((st, g, slot, v))
    frames.append(b"\x01\x02")                        # shorter than a frame header
    want.append((O.WIRE_MALFORMED, 0, 0, None))
    frames.append(np.array([n_groups << 4], dtype="<u4").tobytes() + bytes.fromhex("0804"))   # group out of range
    want.append((O.WIRE_MALFORMED, 0, 0, None))
    blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
    offsets = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.uint32)
    status, recs, hint, snap = O.wire_decode_batch(blob, offsets, n_groups, terms)
    assert [int(s) for s in status] == [w[0] for w in want]
    kinds = set(int(s) for s in status)
    assert kinds == {O.WIRE_OK, O.WIRE_SKIP_TYPE, O.WIRE_TERM, O.WIRE_NEEDS_LOG, O.WIRE_MALFORMED, O.WIRE_DUP}
    for i, (st, g, slot, v) in enumerate(want):
        if st != O.WIRE_OK:
            continue
        f = v["fields"]
        assert (recs["group"][i], recs["peer_slot"][i], recs["flags"][i]) == (g, slot, 1 if f["reject"] else 0)
        assert (int(recs["index"][i]), int(recs["commit"][i])) == (f["index"], f["commit"])
        assert (int(hint[i]), int(snap[i])) == (f["reject_hint"], f["request_snapshot"])
