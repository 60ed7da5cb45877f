// test_mirror.cpp -- the reference's own tracker / quorum / commit tests, restated on the C++
// mirror (raftgpu.hpp) so they read like the originals.  Every assertion runs through
// libraftgpu.so on the GPU.  Run by tests/test_gpu_mirror.py (pytest -m gpu).
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <string>
#include <vector>

#include "raftgpu.hpp"

using namespace raft;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond, ...)                                                  \
    do {                                                                  \
        g_checks++;                                                       \
        if (!(cond)) {                                                    \
            g_failed++;                                                   \
            std::printf("FAIL %s:%d: %s -- ", __FILE__, __LINE__, #cond); \
            std::printf(__VA_ARGS__);                                     \
            std::printf("\n");                                            \
        }                                                                 \
    } while (0)

static std::shared_ptr<Arena> g_arena;

// a tracker holding peers 1..n as voters (incoming), self = 1
struct Fixture {
    ProgressTracker prs;
    LeaderLog log;
    explicit Fixture(size_t n, uint64_t next_idx = 1) : prs(g_arena, 256), log(g_arena, prs) {
        prs.set_self(1);
        Configuration conf;
        MapChange changes;
        std::set<uint64_t> ids;
        for (uint64_t id = 1; id <= n; id++) {
            ids.insert(id);
            changes.emplace_back(id, MapChangeType::Add);
        }
        conf.voters = JointConfig(ids);
        prs.apply_conf(conf, changes, next_idx);
    }
    // progress.rs:250-262 new_progress(state, matched, next_idx, pending_snapshot, ins_size)
    ProgressRef new_progress(ProgressState state, uint64_t matched, uint64_t next_idx, uint64_t pending_snapshot = 0) {
        ProgressRef pr = *prs.get_mut(2);
        Progress p;  // Progress::new(next_idx, ins_size), progress.rs:60-73
        p.next_idx = next_idx;
        p.state = state;
        p.matched = matched;
        p.pending_snapshot = pending_snapshot;
        pr.store(p);
        return pr;
    }
};

// progress.rs:264-283 test_progress_is_paused
static void test_progress_is_paused() {
    struct Row { ProgressState state; bool paused; bool w; };
    const Row tests[] = {{ProgressState::Probe, false, false},    {ProgressState::Probe, true, true},
                         {ProgressState::Replicate, false, false}, {ProgressState::Replicate, true, false},
                         {ProgressState::Snapshot, false, true},   {ProgressState::Snapshot, true, true}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(2);
        ProgressRef pr = f.new_progress(t.state, 0, 0);
        Progress p = pr.load();
        p.paused = t.paused;
        pr.store(p);
        CHECK(pr.is_paused() == t.w, "#%d: shouldwait = %d, want %d", i, pr.is_paused(), t.w);
        i++;
    }
}

// progress.rs:285-295 test_progress_resume
static void test_progress_resume() {
    Fixture f(2);
    ProgressRef pr = f.new_progress(ProgressState::Probe, 0, 2);
    pr.pause();
    pr.maybe_decr_to(1, 1, INVALID_INDEX);
    CHECK(!pr.load().paused, "paused= true, want false");
    pr.pause();
    pr.maybe_update(2);
    CHECK(!pr.load().paused, "paused= true, want false");
}

// progress.rs:297-330 test_progress_become_probe
static void test_progress_become_probe() {
    struct Row { ProgressState st; uint64_t next, pending, wnext; };
    const Row tests[] = {{ProgressState::Replicate, 5, 0, 2}, {ProgressState::Snapshot, 5, 10, 11}, {ProgressState::Snapshot, 5, 0, 2}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(2);
        ProgressRef pr = f.new_progress(t.st, 1, t.next, t.pending);
        pr.become_probe();
        const Progress p = pr.load();
        CHECK(p.state == ProgressState::Probe, "#%d: state", i);
        CHECK(p.matched == 1, "#%d: match = %llu", i, (unsigned long long)p.matched);
        CHECK(p.next_idx == t.wnext, "#%d: next = %llu, want %llu", i, (unsigned long long)p.next_idx, (unsigned long long)t.wnext);
        i++;
    }
}

// progress.rs:332-349 test_progress_become_replicate / become_snapshot
static void test_progress_become_replicate_snapshot() {
    {
        Fixture f(2);
        ProgressRef pr = f.new_progress(ProgressState::Probe, 1, 5);
        pr.become_replicate();
        const Progress p = pr.load();
        CHECK(p.state == ProgressState::Replicate && p.matched == 1 && p.matched + 1 == p.next_idx, "become_replicate");
    }
    {
        Fixture f(2);
        ProgressRef pr = f.new_progress(ProgressState::Probe, 1, 5);
        pr.become_snapshot(10);
        const Progress p = pr.load();
        CHECK(p.state == ProgressState::Snapshot && p.matched == 1 && p.pending_snapshot == 10, "become_snapshot");
    }
}

// progress.rs:351-373 test_progress_update
static void test_progress_update() {
    const uint64_t prev_m = 3, prev_n = 5;
    struct Row { uint64_t update, wm, wn; bool wok; };
    const Row tests[] = {{prev_m - 1, prev_m, prev_n, false}, {prev_m, prev_m, prev_n, false},
                         {prev_m + 1, prev_m + 1, prev_n, true}, {prev_m + 2, prev_m + 2, prev_n + 1, true}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(2);
        ProgressRef pr = f.new_progress(ProgressState::Probe, prev_m, prev_n);
        const bool ok = pr.maybe_update(t.update);
        const Progress p = pr.load();
        CHECK(ok == t.wok, "#%d: ok= %d, want %d", i, ok, t.wok);
        CHECK(p.matched == t.wm, "#%d: match= %llu, want %llu", i, (unsigned long long)p.matched, (unsigned long long)t.wm);
        CHECK(p.next_idx == t.wn, "#%d: next= %llu, want %llu", i, (unsigned long long)p.next_idx, (unsigned long long)t.wn);
        i++;
    }
}

// progress.rs:375-412 test_progress_maybe_decr
static void test_progress_maybe_decr() {
    struct Row { ProgressState state; uint64_t m, n, rejected, last; bool w; uint64_t wn; };
    const Row tests[] = {
        {ProgressState::Replicate, 5, 10, 5, 5, false, 10}, {ProgressState::Replicate, 5, 10, 4, 4, false, 10},
        {ProgressState::Replicate, 5, 10, 9, 9, true, 6},   {ProgressState::Probe, 0, 0, 0, 0, false, 0},
        {ProgressState::Probe, 0, 10, 5, 5, false, 10},     {ProgressState::Probe, 0, 10, 9, 9, true, 9},
        {ProgressState::Probe, 0, 2, 1, 1, true, 1},        {ProgressState::Probe, 0, 1, 0, 0, true, 1},
        {ProgressState::Probe, 0, 10, 9, 2, true, 3},       {ProgressState::Probe, 0, 10, 9, 0, true, 1}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(2);
        ProgressRef pr = f.new_progress(t.state, t.m, t.n);
        const bool got = pr.maybe_decr_to(t.rejected, t.last, 0);
        const Progress p = pr.load();
        CHECK(got == t.w, "#%d: maybeDecrTo= %d, want %d", i, got, t.w);
        CHECK(p.matched == t.m, "#%d: match= %llu, want %llu", i, (unsigned long long)p.matched, (unsigned long long)t.m);
        CHECK(p.next_idx == t.wn, "#%d: next= %llu, want %llu", i, (unsigned long long)p.next_idx, (unsigned long long)t.wn);
        i++;
    }
}

// majority.rs:66-68 doc examples + src/quorum/testdata spot checks (the full files are replayed
// through the C-ABI by tests/test_gpu_parity.py)
static void test_quorum_functions() {
    Arena &a = *g_arena;
    {
        AckIndexer l;
        const uint64_t m[] = {2, 2, 2, 4, 5};
        for (uint64_t i = 0; i < 5; i++) l[i + 1] = Index{m[i], 0};
        const auto r = MajorityConfig({1, 2, 3, 4, 5}).committed_index(a, false, l);
        CHECK(r.first == 2 && !r.second, "[2,2,2,4,5] -> %llu", (unsigned long long)r.first);
    }
    {
        AckIndexer l{{1, {1, 1}}, {2, {2, 2}}, {3, {3, 2}}};
        const auto r = MajorityConfig({1, 2, 3}).committed_index(a, true, l);
        CHECK(r.first == 1 && r.second, "group commit doc example -> %llu", (unsigned long long)r.first);
    }
    {
        const auto r = MajorityConfig().committed_index(a, false, {});  // majority.rs:71-75
        CHECK(r.first == UINT64_MAX && r.second, "empty config");
    }
    {   // joint_commit.txt: cfg=(1,3) cfgj=(2) idx=(100,45,50) [ids in order 1,3,2] -> 45
        AckIndexer l{{1, {100, 0}}, {3, {45, 0}}, {2, {50, 0}}};
        const auto r = JointConfig(MajorityConfig({1, 3}), MajorityConfig({2})).committed_index(a, false, l);
        CHECK(r.first == 45, "joint (1,3)x(2) -> %llu", (unsigned long long)r.first);
        const auto s = JointConfig(MajorityConfig({2}), MajorityConfig({1, 3})).committed_index(a, false, l);
        CHECK(s == r, "symmetry");
    }
    {   // joint_group_commit.txt: cfg=(1,2,3,4) cfgj=(3,4,5,6) idx=(101,99,100,102,103,1) gid=(1,_,1,1,_,2) -> 1
        AckIndexer l{{1, {101, 1}}, {2, {99, 0}}, {3, {100, 1}}, {4, {102, 1}}, {5, {103, 0}}, {6, {1, 2}}};
        const auto r = JointConfig(MajorityConfig({1, 2, 3, 4}), MajorityConfig({3, 4, 5, 6})).committed_index(a, true, l);
        CHECK(r.first == 1, "joint group commit -> %llu", (unsigned long long)r.first);
    }
    {   // majority_vote.txt / joint_vote.txt
        auto votes = [](std::map<uint64_t, bool> m) {
            return [m](uint64_t id) -> std::optional<bool> {
                const auto it = m.find(id);
                return it == m.end() ? std::nullopt : std::optional<bool>(it->second);
            };
        };
        CHECK(MajorityConfig().vote_result(a, votes({})) == VoteResult::Won, "empty config wins");
        CHECK(MajorityConfig({4, 8}).vote_result(a, votes({{4, false}})) == VoteResult::Lost, "(4,8) n,_ loses");
        CHECK(MajorityConfig({4, 8}).vote_result(a, votes({{4, true}})) == VoteResult::Pending, "(4,8) y,_ pending");
        CHECK(JointConfig(MajorityConfig({1}), MajorityConfig({2})).vote_result(a, votes({{1, true}, {2, true}})) == VoteResult::Won, "(1)x(2) y,y");
        CHECK(JointConfig(MajorityConfig({1}), MajorityConfig({2})).vote_result(a, votes({{1, true}, {2, false}})) == VoteResult::Lost, "(1)x(2) y,n");
    }
}

// harness/tests/integration_cases/test_raft.rs:1145-1240 test_commit
static void test_commit() {
    struct Row { std::vector<uint64_t> matches; std::vector<std::pair<uint64_t, uint64_t>> logs; uint64_t sm_term, w; };
    const std::vector<Row> tests = {
        {{1}, {{1, 1}}, 1, 1}, {{1}, {{1, 1}}, 2, 0}, {{2}, {{1, 1}, {2, 2}}, 2, 2}, {{1}, {{2, 1}}, 2, 1},
        {{2, 1, 1}, {{1, 1}, {2, 2}}, 1, 1}, {{2, 1, 1}, {{1, 1}, {1, 2}}, 2, 0},
        {{2, 1, 2}, {{1, 1}, {2, 2}}, 2, 2}, {{2, 1, 2}, {{1, 1}, {1, 2}}, 2, 0},
        {{2, 1, 1, 1}, {{1, 1}, {2, 2}}, 1, 1}, {{2, 1, 1, 1}, {{1, 1}, {1, 2}}, 2, 0},
        {{2, 1, 1, 2}, {{1, 1}, {2, 2}}, 1, 1}, {{2, 1, 1, 2}, {{1, 1}, {1, 2}}, 2, 0},
        {{2, 1, 2, 2}, {{1, 1}, {2, 2}}, 2, 2}, {{2, 1, 2, 2}, {{1, 1}, {1, 2}}, 2, 0}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(t.matches.size());
        for (size_t j = 0; j < t.matches.size(); j++) {
            ProgressRef pr = *f.prs.get_mut(j + 1);
            Progress p = pr.load();
            p.matched = t.matches[j];      // pr.matched = *v; pr.next_idx = *v + 1  (test_raft.rs:1230-1232)
            p.next_idx = t.matches[j] + 1;
            pr.store(p);
        }
        // the log: entries (term, index); the leader's term is sm_term
        uint64_t term_start = RAFTGPU_NO_TERM_START;
        for (const auto &[term, index] : t.logs)
            if (term == t.sm_term && index < term_start) term_start = index;
        f.log.set_log_bounds(term_start, t.logs.back().second);
        f.log.maybe_commit();
        CHECK(f.log.committed() == t.w, "#%d: committed = %llu, want %llu", i, (unsigned long long)f.log.committed(), (unsigned long long)t.w);
        i++;
    }
}

// test_raft.rs:5092-5163 test_group_commit
static void test_group_commit() {
    struct Row { std::vector<uint64_t> matches, gids; uint64_t g_w, q_w; };
    const std::vector<Row> tests = {
        {{1}, {0}, 1, 1}, {{1}, {1}, 1, 1},
        {{2, 2, 1}, {1, 2, 1}, 2, 2}, {{2, 2, 1}, {1, 1, 2}, 1, 2}, {{2, 2, 1}, {1, 0, 1}, 1, 2}, {{2, 2, 1}, {0, 0, 0}, 1, 2},
        {{4, 2, 1, 3}, {0, 0, 0, 0}, 1, 2}, {{4, 2, 1, 3}, {1, 0, 0, 0}, 1, 2}, {{4, 2, 1, 3}, {0, 1, 0, 2}, 2, 2},
        {{4, 2, 1, 3}, {0, 2, 1, 0}, 1, 2}, {{4, 2, 1, 3}, {1, 1, 1, 1}, 2, 2}, {{4, 2, 1, 3}, {1, 1, 2, 1}, 1, 2},
        {{4, 2, 1, 3}, {1, 2, 1, 1}, 2, 2}, {{4, 2, 1, 3}, {4, 3, 2, 1}, 2, 2}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(t.matches.size());
        uint64_t lo = UINT64_MAX, hi = 0;
        for (size_t j = 0; j < t.matches.size(); j++) {
            ProgressRef pr = *f.prs.get_mut(j + 1);
            Progress p = pr.load();
            p.matched = t.matches[j];
            p.next_idx = t.matches[j] + 1;
            p.commit_group_id = t.gids[j];       // assign_commit_groups, raft.rs:531-540
            pr.store(p);
            lo = std::min(lo, t.matches[j]);
            hi = std::max(hi, t.matches[j]);
        }
        f.log.set_log_bounds(lo, hi);             // logs = (min..=max) all of term 1 = the leader's
        f.prs.enable_group_commit(true);
        f.log.maybe_commit();
        CHECK(f.log.committed() == t.g_w, "#%d: leader group committed %llu, want %llu", i, (unsigned long long)f.log.committed(), (unsigned long long)t.g_w);
        f.prs.enable_group_commit(false);         // raft.rs:513-518: disabling re-runs maybe_commit
        f.log.maybe_commit();
        CHECK(f.log.committed() == t.q_w, "#%d: quorum committed %llu, want %llu", i, (unsigned long long)f.log.committed(), (unsigned long long)t.q_w);
        i++;
    }
}

// test_raft.rs:2611-2675 test_leader_append_response, through the batched driver
static void test_leader_append_response() {
    struct Row { uint64_t index; bool reject; uint64_t wmatch, wnext; uint64_t wcommitted; };
    // wnext is the value before the send path's optimistic_update (the reference's 4 in row 3
    // includes bcast_append -> update_state, which stays with the caller)
    const Row tests[] = {{3, true, 0, 3, 0}, {2, true, 0, 2, 0}, {2, false, 2, 3, 2}, {0, false, 0, 3, 0}};
    int i = 0;
    for (const Row &t : tests) {
        Fixture f(3);
        // log terms [0, 1] persisted, leader elected at term 1: reset(last=2, committed=0, persisted=2),
        // become_leader appends the noop at 3 (term_start = 3)... the entries of term 1 are 2 and 3
        f.log.reset(2, 0, 2);
        f.log.become_leader();
        f.log.set_log_bounds(2, 3);
        for (uint64_t id : {2, 3}) f.prs.get_mut(id)->pause();  // bcast_append in Probe pauses the peer
        MultiRaftDriver drv(g_arena);
        drv.step_append_response(f.prs, 2, t.index, 0, t.reject, t.index);
        drv.step();
        const Progress p = *f.prs.get(2);
        CHECK(p.matched == t.wmatch, "#%d: match = %llu, want %llu", i, (unsigned long long)p.matched, (unsigned long long)t.wmatch);
        CHECK(p.next_idx == t.wnext, "#%d: next = %llu, want %llu", i, (unsigned long long)p.next_idx, (unsigned long long)t.wnext);
        CHECK(f.log.committed() == t.wcommitted, "#%d: commit = %llu, want %llu", i, (unsigned long long)f.log.committed(), (unsigned long long)t.wcommitted);
        CHECK(p.recent_active, "#%d: recent_active", i);
        i++;
    }
}

// A whole tick as one call, then the post-commit send decisions: the acknowledgement that commits
// index 3 makes the leader bcast_append (raft.rs:1745-1748 -> 857-865) to every follower that is not
// paused (raft.rs:780-788); test_raft_paper.rs:499-534 test_leader_commit_entry expects exactly
// those MsgAppends.  Two responses of one peer in one tick apply in arrival order.
static void test_tick_batch_and_send_list() {
    Fixture f(3);
    f.log.reset(2, 0, 2);
    f.log.become_leader();
    f.log.set_log_bounds(2, 3);
    for (uint64_t id : {2, 3}) f.prs.get_mut(id)->become_replicate();
    MultiRaftDriver drv(g_arena);
    drv.push_append_response(f.prs, 2, 2, 0);
    drv.push_append_response(f.prs, 2, 3, 2);   // the same peer again: pipelined acks
    drv.push_local_progress(f.prs, 1, 3, 3);    // the leader has persisted its own entry
    const raftgpu_step_result r = drv.step_tick();
    CHECK(r.n_records == 3 && r.n_advanced == 1 && r.n_duplicates == 0, "tick: %llu records, %llu advanced",
          (unsigned long long)r.n_records, (unsigned long long)r.n_advanced);
    CHECK(f.log.committed() == 3 && drv.advanced(f.prs.group()), "committed = %llu, want 3", (unsigned long long)f.log.committed());
    CHECK(f.prs.get(2)->matched == 3 && f.prs.get(2)->next_idx == 4 && f.prs.get(2)->committed_index == 2, "peer 2 after both acks");
    auto sends = drv.send_list();
    CHECK(sends.size() == 2, "send list has %zu entries, want 2 (both followers)", sends.size());
    std::set<uint64_t> to;
    for (const auto &e : sends) {
        CHECK(e.group == f.prs.group() && e.flags == 0, "entry of another group / unexpected flags");
        to.insert(f.prs.id_of(e.peer_slot));
        CHECK(e.next_idx == f.prs.get(f.prs.id_of(e.peer_slot))->next_idx, "entry carries the peer's next_idx");
    }
    CHECK(to == (std::set<uint64_t>{2, 3}), "MsgAppend goes to 2 and 3, never to the leader itself");
    // a paused probe and a full inflight window are skipped (progress.rs:210-216)
    f.prs.get_mut(3)->become_probe();
    f.prs.get_mut(3)->pause();
    f.log.set_log_bounds(2, 4);
    drv.push_append_response(f.prs, 2, 4, 3);
    drv.push_local_progress(f.prs, 1, 4, 4);
    drv.step_tick();
    sends = drv.send_list();
    CHECK(f.log.committed() == 4 && sends.size() == 1 && f.prs.id_of(sends[0].peer_slot) == 2, "only the unpaused follower is sent to");
    // nothing advanced -> nothing to broadcast
    drv.push_append_response(f.prs, 2, 4, 4);
    drv.step_tick();
    CHECK(drv.send_list().empty(), "no commit, no bcast_append");
}

// check_quorum (raft.rs:1963-1973) -> quorum_recently_active / has_quorum (tracker.rs:346-372)
static void test_quorum_activity() {
    Fixture f(3);
    // apply_conf marks new peers recently active (tracker.rs:385-389)
    CHECK(f.prs.quorum_recently_active(1), "all three were just added");
    CHECK(!f.prs.get(2)->recent_active && !f.prs.get(3)->recent_active && f.prs.get(1)->recent_active, "flags cleared except self");
    CHECK(!f.prs.quorum_recently_active(1), "nobody spoke since the last check");
    Progress p = *f.prs.get(3);
    p.recent_active = true;
    f.prs.get_mut(3)->store(p);
    CHECK(f.prs.quorum_recently_active(1), "self + 3 = quorum of 3");
    CHECK(f.prs.has_quorum({1, 2}) && !f.prs.has_quorum({3}) && !f.prs.has_quorum({7, 8, 9}), "has_quorum");
    // votes: tracker.rs:301-340
    f.prs.reset_votes();
    f.prs.record_vote(1, true);
    f.prs.record_vote(2, false);
    f.prs.record_vote(2, true);  // or_insert: the first vote stands
    auto [granted, rejected, res] = f.prs.tally_votes();
    CHECK(granted == 1 && rejected == 1 && res == VoteResult::Pending, "tally %zu/%zu", granted, rejected);
    f.prs.record_vote(3, true);
    CHECK(std::get<2>(f.prs.tally_votes()) == VoteResult::Won, "won with 1 and 3");
    CHECK(f.prs.vote_result({{1, false}, {2, false}}) == VoteResult::Lost, "explicit map");
}

// error behaviour: StepPeerNotFound (raw_node.rs:402-411), commit_to fatal! (raft_log.rs:291-298),
// update_state panic (progress.rs:238-241)
static void test_errors() {
    Fixture f(3);
    MultiRaftDriver drv(g_arena);
    bool threw = false;
    try { drv.step_append_response(f.prs, 9, 1, 0); } catch (const StepPeerNotFound &) { threw = true; }
    CHECK(threw, "unknown responder -> StepPeerNotFound");
    f.log.reset(3, 2, 3);
    f.log.commit_to(3);
    CHECK(f.log.committed() == 3, "commit_to 3");
    f.log.commit_to(1);
    CHECK(f.log.committed() == 3, "never decrease");  // raft_log.rs:1497-1522 test_commit_to
    threw = false;
    try { f.log.commit_to(4); } catch (const Fatal &) { threw = true; }
    CHECK(threw, "commit out of range -> fatal");
    threw = false;
    ProgressRef pr = *f.prs.get_mut(2);
    pr.become_snapshot(5);
    try { pr.update_state(9); } catch (const Fatal &) { threw = true; }
    CHECK(threw, "update_state in Snapshot panics");
    CHECK(!f.prs.get(42).has_value(), "get(unknown) is None");
}

// More than 8 peers: the reference sorts any number of voters (majority.rs:86-93); here the group is a WIDE one.
// A joint change between two disjoint 5-voter sets (ids 1..5 -> 6..10) plus a learner (11), leader = 1: an index is
// committed only when a majority of BOTH sets holds it (joint.rs:47-51).
static void test_wide_group() {
    Arena &a = *g_arena;
    {   // free-standing quorum functions over 11 and 13 ids (a wide scratch group)
        AckIndexer l;
        for (uint64_t id = 1; id <= 11; id++) l[id] = Index{id * 10, 0};
        const auto r = MajorityConfig({1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}).committed_index(a, false, l);
        CHECK(r.first == 60, "11 voters [10..110] -> %llu, want 60", (unsigned long long)r.first);   // the 6th largest
        const auto j = JointConfig(MajorityConfig({1, 2, 3, 4, 5, 6, 7}), MajorityConfig({8, 9, 10, 11, 12, 13})).committed_index(a, false, l);
        // incoming: 4th largest of [70..10] = 40; outgoing (12 and 13 never acked): 4th largest of [110,100,90,80,0,0] = 80
        CHECK(j.first == 40, "joint (1..7)x(8..13) -> %llu, want 40", (unsigned long long)j.first);
    }
    ProgressTracker prs(g_arena, 256, /*wide=*/true);
    LeaderLog log(g_arena, prs);
    prs.set_self(1);
    Configuration conf;
    MapChange changes;
    std::set<uint64_t> in, out;
    for (uint64_t id = 1; id <= 11; id++) {
        changes.emplace_back(id, MapChangeType::Add);
        if (id <= 5) in.insert(id);
        else if (id <= 10) out.insert(id);
    }
    conf.voters = JointConfig(MajorityConfig(in), MajorityConfig(out));
    conf.learners.insert(11);
    prs.apply_conf(conf, changes, 1);
    log.reset(2, 0, 2);
    log.become_leader();
    log.set_log_bounds(2, 3);
    for (uint64_t id = 2; id <= 11; id++) prs.get_mut(id)->become_replicate();
    MultiRaftDriver drv(g_arena);
    drv.push_local_progress(prs, 1, 3, 3);
    for (uint64_t id : {2, 3, 6, 7}) drv.push_append_response(prs, id, 3, 0);
    drv.push_append_response(prs, 11, 3, 0);  // the learner's ack does not count
    raftgpu_step_result r = drv.step_tick();
    CHECK(r.n_advanced == 0 && log.committed() == 0, "3 of the incoming set but 2 of the outgoing: nothing commits (committed = %llu)", (unsigned long long)log.committed());
    CHECK(prs.get(7)->matched == 3 && prs.get(11)->matched == 3, "acks of peers in the high half land");
    drv.push_append_response(prs, 9, 3, 0);   // peer slot 8: the first cell of the high half
    r = drv.step_tick();
    CHECK(r.n_advanced == 1 && log.committed() == 3 && drv.advanced(prs.group()), "third ack of the outgoing set commits 3 (committed = %llu)", (unsigned long long)log.committed());
    CHECK(prs.maximal_committed_index().first == 3, "maximal_committed_index over 10 voters");
    CHECK(prs.has_quorum({1, 2, 3, 6, 7, 9}) && !prs.has_quorum({1, 2, 3, 4, 5, 6, 7}), "has_quorum needs both majorities");
    bool threw = false;
    try {
        MapChange more;
        Configuration c2 = conf;
        for (uint64_t id = 12; id <= 17; id++) {
            more.emplace_back(id, MapChangeType::Add);
            c2.learners.insert(id);
        }
        prs.apply_conf(c2, more, 1);
    } catch (const Error &e) { threw = e.status == RAFTGPU_ERR_TOO_MANY_PEERS; }
    CHECK(threw, "a 17th peer -> RAFTGPU_ERR_TOO_MANY_PEERS");
}

int main() {
    try {
        g_arena = Arena::create(0, 4096);
    } catch (const Error &e) {
        std::printf("cannot create arena: %s\n", e.what());
        return 2;
    }
    test_progress_is_paused();
    test_progress_resume();
    test_progress_become_probe();
    test_progress_become_replicate_snapshot();
    test_progress_update();
    test_progress_maybe_decr();
    test_quorum_functions();
    test_commit();
    test_group_commit();
    test_leader_append_response();
    test_tick_batch_and_send_list();
    test_quorum_activity();
    test_errors();
    test_wide_group();
    std::printf("%s: %d checks, %d failed\n", g_failed ? "FAILED" : "ok", g_checks, g_failed);
    g_arena.reset();
    return g_failed ? 1 : 0;
}
