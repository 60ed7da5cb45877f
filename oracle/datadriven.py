"""Replay of the reference's data-driven quorum goldens against the C oracle.

TEST INFRASTRUCTURE ONLY.  Re-implements, for the five files under
``src/quorum/testdata/`` (committed as tests/golden/quorum/), what the
reference's runner does:

* file format   -- datadriven/src/test_data_reader.rs:37-137, 140-205 and
                   datadriven/src/line_sparser.rs:19-69
* the commands  -- src/quorum/datadriven_test.rs:5-312 (``committed``,
                   ``group_committed``, ``vote``), including its metamorphic
                   checks (symmetry :177-181, zero-/self-joint :187-199,
                   overlay :204-245)
* ``describe``  -- src/quorum/majority.rs:170-238, src/quorum/joint.rs:92-97,
                   ``Index`` Display src/quorum.rs:40-55

The arithmetic itself is done by oracle/raft_oracle.c through ``committed`` /
``vote`` callables, so the same replay can also be pointed at another
implementation (the GPU path) to pin it against the same vectors.
"""
from __future__ import annotations

import dataclasses
import re

from . import oracle as O

U64_MAX = O.U64_MAX


@dataclasses.dataclass
class TestData:
    pos: str
    cmd: str
    cmd_args: list[tuple[str, list[str]]]
    input: str
    expected: str


_DIRECTIVE = re.compile(r"^ *[-a-zA-Z0-9/_,.]+(|=[-a-zA-Z0-9_@=+/,.]*|=\([^)]*\))( |$)")


def _split_directives(line: str) -> list[str]:
    # line_sparser.rs:71-95: repeatedly match the directive regex from the left
    res = []
    rest = line
    while rest:
        m = _DIRECTIVE.match(rest)
        if not m:
            raise ValueError(f"cannot parse directive at column {len(line) - len(rest)}: {line!r}")
        res.append(m.group(0).strip())
        rest = rest[m.end():]
    return res


def parse_line(line: str):
    fields = _split_directives(line)
    if not fields:
        return "", []
    cmd, args = fields[0], []
    for arg in fields[1:]:
        kv = arg.split("=", 1)
        if len(kv) == 1:
            args.append((kv[0], []))
        else:
            key, val = kv
            if val.startswith("(") and val.endswith(")"):
                args.append((key, [v.strip() for v in val[1:-1].split(",")]))
            else:
                args.append((key, [val]))
    return cmd, args


def parse_file(path: str) -> list[TestData]:
    with open(path, encoding="utf-8") as f:
        lines = f.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    out, i, n = [], 0, len(lines)
    while i < n:
        line = lines[i].strip()
        pos = i + 1
        i += 1
        if line.startswith("#") or not line:
            continue
        while line.endswith("\\"):
            nxt = lines[i].strip()
            i += 1
            line = line[:-1]
            if nxt:
                line += " " + nxt
        cmd, args = parse_line(line)
        if not cmd:
            raise ValueError("cmd must not be empty")
        buf, separator = [], False
        while i < n:
            l = lines[i]
            i += 1
            if l == "----":
                separator = True
                break
            buf.append(l)
        expected = ""
        if separator and i < n:
            if lines[i] == "----":
                # double-separator form: blank lines allowed until "----\n----"
                i += 1
                while True:
                    l = lines[i]
                    i += 1
                    if l == "----" and lines[i] == "----":
                        i += 1
                        if i < n:
                            assert lines[i] == ""
                            i += 1
                        break
                    expected += l + "\n"
            else:
                while i < n and lines[i].strip():
                    expected += lines[i] + "\n"
                    i += 1
        out.append(TestData(f"{path} : L{pos}", cmd, args, "\n".join(buf).strip(), expected))
    return out


# ------------------------------------------------------------------ Display / describe

def fmt_index(index: int, group_id: int = 0) -> str:
    s = "∞" if index == U64_MAX else str(index)
    return s if group_id == 0 else f"[{group_id}]{s}"


def describe(voters: list[int], lookup: dict[int, tuple[int, int]]) -> str:
    n = len(voters)
    if n == 0:
        return "<empty majority quorum>"
    info = [{"id": v, "idx": lookup.get(v), "bar": 0} for v in voters]
    key = lambda t: ((t["idx"] or (0, 0))[0], t["id"])
    info.sort(key=key)
    for i in range(1, n):
        if (info[i - 1]["idx"] or (0, 0))[0] < (info[i]["idx"] or (0, 0))[0]:
            info[i]["bar"] = i
    info.sort(key=lambda t: t["id"])
    buf = " " * n + "    idx\n"
    for t in info:
        if t["idx"] is not None:
            buf += "x" * t["bar"] + ">" + " " * (n - t["bar"])
            buf += " {:>5}    (id={})\n".format(fmt_index(*t["idx"]), t["id"])
        else:
            buf += "?" + " " * n
            buf += " {:>5}    (id={})\n".format(fmt_index(0, 0), t["id"])
    return buf


# ------------------------------------------------------------------ the command runner

def _default_majority(voters, lookup, use_gc):
    return O.majority_committed_index(voters, lookup, use_gc)


def _default_joint(incoming, outgoing, lookup, use_gc):
    return O.joint_committed_index(incoming, outgoing, lookup, use_gc)


def _default_majority_vote(voters, votes):
    return O.majority_vote_result(voters, votes)


def _default_joint_vote(incoming, outgoing, votes):
    return O.joint_vote_result(incoming, outgoing, votes)


class Impl:
    """The four entry points a replay exercises; defaults to the C oracle."""

    def __init__(self, majority=_default_majority, joint=_default_joint,
                 majority_vote=_default_majority_vote, joint_vote=_default_joint_vote):
        self.majority, self.joint = majority, joint
        self.majority_vote, self.joint_vote = majority_vote, joint_vote


def _dedup(seq):
    seen, out = set(), []
    for x in seq:
        if x not in seen:
            seen.add(x)
            out.append(x)
    return out


def run_quorum_case(d: TestData, impl: Impl | None = None) -> str:
    impl = impl or Impl()
    joint = False
    ids, idsj, idxs, gids, votes = [], [], [], [], []
    for key, vals in d.cmd_args:
        for val in vals:
            if key == "cfg":
                ids.append(int(val))
            elif key == "cfgj":
                joint = True
                if val == "zero":
                    assert len(vals) == 1, "cannot mix 'zero' into configuration"
                else:
                    idsj.append(int(val))
            elif key == "idx":
                n = 0
                if val != "_":
                    n = int(val)
                    assert n != 0, "use '_' as 0"
                idxs.append([n, 0])
            elif key == "gid":
                n = 0
                if val != "_":
                    n = int(val)
                    assert n != 0, "use '_' as 0"
                gids.append(n)
            elif key == "votes":
                votes.append([{"y": 2, "n": 1, "_": 0}[val], 0])
            else:
                raise ValueError(f"unknown arg: {key}")

    c, cj = _dedup(ids), _dedup(idsj)  # HashSet<u64>

    def make_lookuper(vals):
        l, p = {}, 0
        for i in ids + idsj:
            if i not in l and p < len(vals):
                l[i] = (vals[p][0], vals[p][1])
                p += 1
        return {k: v for k, v in l.items() if v[0] > 0}

    inp = len(votes) if d.cmd == "vote" else len(idxs)
    voters = len(set(c) | set(cj))
    if voters != inp:
        return f"error: mismatched input (explicit or _) for voters {voters}: {inp}"
    if gids:
        if len(gids) != voters:
            return f"error: mismatched input (explicit or _) for group ids {voters}: {len(gids)}"
        for ix, g in zip(idxs, gids):
            ix[1] = g

    buf = ""
    if d.cmd == "committed":
        l = make_lookuper(idxs)
        if joint:
            buf += describe(_dedup(c + cj), l)
            idx = impl.joint(c, cj, l, False)
            a_idx = impl.joint(cj, c, l, False)
            if a_idx != idx:
                buf += f"{a_idx[0]} <-- via symmetry\n"
        else:
            idx = impl.majority(c, l, False)
            buf += describe(c, l)
            a_idx = impl.joint(c, [], l, False)
            if a_idx != idx:
                buf += f"{a_idx[0]} <-- via zero-joint quorum\n"
            a_idx = impl.joint(c, c, l, False)
            if a_idx != idx:
                buf += f"{a_idx[0]} <-- via self-joint quorum\n"
            for i in c:
                if i in l:
                    iidx = l[i]
                    if idx[0] > iidx[0]:
                        l[i] = (iidx[0] - 1, iidx[1])
                        a_idx = impl.majority(c, l, False)
                        if a_idx != idx:
                            buf += f"{a_idx[0]} <-- overlaying {i}->{iidx[0] - 1}\n"
                        l[i] = (0, iidx[1])
                        a_idx = impl.majority(c, l, False)
                        if a_idx != idx:
                            buf += f"{a_idx[0]} <-- overlaying {i}->0\n"
                        l[i] = iidx
        buf += fmt_index(idx[0]) + "\n"
    elif d.cmd == "group_committed":
        l = make_lookuper(idxs)
        idx = (0, False)
        if joint:
            idx = impl.joint(c, cj, l, True)
            a_idx = impl.joint(cj, c, l, True)
            if a_idx != idx:
                buf += f"{a_idx[0]} <-- via symmetry\n"
        buf += fmt_index(idx[0]) + "\n"
    elif d.cmd == "vote":
        ll = make_lookuper(votes)
        l = {i: (v[0] != 1) for i, v in ll.items()}
        if joint:
            r = impl.joint_vote(c, cj, l)
            ar = impl.joint_vote(cj, c, l)
            if ar != r:
                buf += f"{O.VOTE_NAMES[ar]} <-- via symmetry\n"
        else:
            r = impl.majority_vote(c, l)
        buf += O.VOTE_NAMES[r] + "\n"
    else:
        raise ValueError(f"unknown command: {d.cmd}")
    return buf


def replay_file(path: str, impl: Impl | None = None):
    """Returns [(TestData, actual)] for every directive of one golden file."""
    return [(d, run_quorum_case(d, impl)) for d in parse_file(path)]
