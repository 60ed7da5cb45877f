"""Pins the C oracle against the reference's own golden vectors.

Replays every directive of src/quorum/testdata/*.txt (committed verbatim under
tests/golden/quorum/) the way src/quorum/datadriven_test.rs:5-319 does and
requires FULL-TEXT equality with the text under ``----`` (bar chart of
``describe`` + metamorphic-check lines + result), not just the last line.
"""
import itertools
import os
import random

import pytest

from oracle import datadriven as dd
from oracle import oracle as O

FILES = {
    "majority_commit.txt": 16,
    "joint_commit.txt": 50,
    "joint_group_commit.txt": 14,
    "majority_vote.txt": 22,
    "joint_vote.txt": 39,
}


@pytest.mark.parametrize("name,count", sorted(FILES.items()))
def test_golden_file_full_text(golden_dir, name, count):
    cases = dd.replay_file(os.path.join(golden_dir, "quorum", name))
    assert len(cases) == count, f"{name}: parsed {len(cases)} directives, expected {count}"
    for d, actual in cases:
        assert actual == d.expected, f"{d.pos}\n--- actual\n{actual}--- expected\n{d.expected}"


def _shuffled_impl(rng):
    """Same oracle, but voter sets are presented in a random iteration order
    (HashSet order in the reference is arbitrary; results must not depend on it)."""
    def sh(x):
        x = list(x)
        rng.shuffle(x)
        return x
    return dd.Impl(
        majority=lambda v, l, gc: O.majority_committed_index(sh(v), l, gc),
        joint=lambda a, b, l, gc: O.joint_committed_index(sh(a), sh(b), l, gc),
        majority_vote=lambda v, votes: O.majority_vote_result(sh(v), votes),
        joint_vote=lambda a, b, votes: O.joint_vote_result(sh(a), sh(b), votes),
    )


@pytest.mark.parametrize("name", sorted(FILES))
def test_golden_independent_of_voter_order(golden_dir, name):
    rng = random.Random(0xC0FFEE)
    path = os.path.join(golden_dir, "quorum", name)
    for _ in range(20):
        for d, actual in dd.replay_file(path, _shuffled_impl(rng)):
            assert actual == d.expected, d.pos


def test_majority_doc_examples():
    # majority.rs:66-68 doc comment
    l = {i + 1: (v, 0) for i, v in enumerate([2, 2, 2, 4, 5])}
    assert O.majority_committed_index([1, 2, 3, 4, 5], l) == (2, False)
    l = {1: (1, 1), 2: (2, 2), 3: (3, 2)}
    assert O.majority_committed_index([1, 2, 3], l, True) == (1, True)
    # majority.rs:71-75
    assert O.majority_committed_index([], {}) == (O.U64_MAX, True)
    for n, q in [(0, 1), (1, 1), (2, 2), (3, 2), (4, 3), (5, 3), (7, 4), (8, 5)]:
        assert O.lib().ro_majority(n) == q  # util.rs:118-120


def test_more_than_seven_voters_heap_path():
    # majority.rs:86-93: > 7 voters takes the Vec path; same answer.
    rng = random.Random(7)
    for n in range(8, 17):
        for _ in range(50):
            vals = [rng.randrange(0, 50) for _ in range(n)]
            l = {i + 1: (v, 0) for i, v in enumerate(vals)}
            want = sorted(vals, reverse=True)[n // 2]
            assert O.majority_committed_index(list(range(1, n + 1)), l) == (want, False)


def _gc_closed_form(vals, gids):
    """order-independent closed form of majority.rs:102-123."""
    n = len(vals)
    q_idx = sorted(vals, reverse=True)[n // 2]
    top = {}
    for v, g in zip(vals, gids):
        if g:
            top[g] = max(top.get(g, 0), v)
    if len(top) >= 2:
        return min(q_idx, sorted(top.values(), reverse=True)[1]), True
    if len(top) == 1 and all(g != 0 for g in gids):
        return q_idx, False
    return min(vals), False


def test_group_commit_closed_form_and_order_independence():
    rng = random.Random(12345)
    for _ in range(4000):
        n = rng.randrange(1, 8)
        vals = [rng.randrange(0, 12) for _ in range(n)]
        gids = [rng.randrange(0, 4) for _ in range(n)]
        ids = list(range(1, n + 1))
        l = {i: (v, g) for i, v, g in zip(ids, vals, gids)}
        want = _gc_closed_form(vals, gids)
        for _ in range(4):
            rng.shuffle(ids)
            assert O.majority_committed_index(ids, l, True) == want, (vals, gids, ids)


def test_joint_properties_random():
    rng = random.Random(99)
    for _ in range(2000):
        universe = list(range(1, 9))
        a = rng.sample(universe, rng.randrange(0, 6))
        b = rng.sample(universe, rng.randrange(0, 6))
        l = {i: (rng.randrange(0, 30), 0) for i in universe if rng.random() < 0.85}
        ia = O.majority_committed_index(a, l)
        ib = O.majority_committed_index(b, l)
        j = O.joint_committed_index(a, b, l)
        assert j[0] == min(ia[0], ib[0])                       # joint.rs:50
        assert j == O.joint_committed_index(b, a, l)            # symmetry
        assert O.joint_committed_index(a, [], l)[0] == ia[0]    # zero-joint
        assert O.joint_committed_index(a, a, l)[0] == ia[0]     # self-joint


def test_quorum_select_header_against_a_sort(tmp_path):
    """raft-rs_b200/csrc/quorum_select.h (the joint quorum index the general kernels use: one comparison pass, two
    masks) built for the HOST and compared with a plain sort over random values with many ties and every kind of mask
    pair (scripts/micro/quorum_select_test.cpp).  The same code runs on the device."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "qs_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "raft-rs_b200", "csrc"),
                    os.path.join(root, "scripts", "micro", "quorum_select_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "600"], capture_output=True, text=True)
    assert out.returncode == 0 and "quorum select ok" in out.stdout, out.stdout + out.stderr
