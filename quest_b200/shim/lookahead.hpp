// lookahead.hpp -- the gate log behind the look-ahead swap-in victim choice (quest_b200/shim/localiser_b200.cpp owns the
// only instance; tests/test_lookahead_cpu.py compiles this header alone and drives it with bench.py's gate streams).
//
// Which shard qubit a swap-in evicts decides how soon the next half-shard exchange comes: the evicted qubit costs
// another exchange the next time a gate needs it inside the shard.  QuEST's API is eager, so at the gate that forces
// the swap-in nothing is known about the gates that follow, and the default rule (least recently used) is blind.
// With a window W > 0 (QUEST_B200_LOOKAHEAD=W; 0 = off = default) the relabelling-aware gate entry points of
// distributed GPU statevectors do not run when called: each call is logged by value, in program order, and the log
// is replayed through the very same entry points
//   - when it holds 2W gates: the older W run, each with at least W gates of known future, or
//   - when anything else touches a Qureg (every other entry point, host copies, syncQuESTEnv, destruction -- the
//     places that also restore the canonical qubit order).
// During the replay the swap-in evicts the candidate whose next use as a non-diagonal target lies farthest ahead in
// the log (Belady's rule over the window); tools/exchange_policy_study.py models the effect on bench.py's circuits.
// Every rank logs and replays the same calls, so the choice is rank-independent like the rest of the relabelling.
// Gates whose matrix lives in user-owned heap memory (CompMatr, DiagMatr, FullStateDiagMatr) are never logged: the
// user may change or free it before the replay.
//
// Reference behaviour this sits in front of: core/localiser.cpp:997-1040 (swap in, apply, swap back, per gate).
#ifndef QB_LOOKAHEAD_HPP
#define QB_LOOKAHEAD_HPP

#include <cstddef>
#include <cstdint>
#include <deque>
#include <functional>
#include <utility>
#include <vector>

namespace qb_lookahead {

struct LoggedGate {
    const void* key;                    // which state (the Qureg's device pointer)
    std::function<void()> run;          // the logged call
    std::vector<int> shardTargs;        // logical qubits it needs inside the shard (its non-diagonal targets)
    int swapA, swapB;                   // an uncontrolled SWAP: these two logical labels trade places (else -1)
};

class GateLog {
public:
    explicit GateLog(int window = 0) : window_(window > 0 ? (size_t) window : 0) {}
    void setWindow(int window) { window_ = window > 0 ? (size_t) window : 0; }
    size_t window() const { return window_; }
    bool replaying() const { return replaying_; }
    bool empty() const { return gates_.empty(); }
    size_t size() const { return gates_.size(); }

    void push(const void* key, std::vector<int> shardTargs, std::function<void()> run, int swapA = -1, int swapB = -1) {
        gates_.push_back(LoggedGate{key, std::move(run), std::move(shardTargs), swapA, swapB});
        if (gates_.size() >= 2 * window_)
            replay(gates_.size() - window_);
    }

    // every path that looks at or changes amplitudes other than through a logged gate calls this first
    void drain() {
        if (!replaying_ && !gates_.empty())
            replay(gates_.size());
    }

    // next[l] = position in the log of the next gate that needs logical qubit l -- as labelled NOW, i.e. at the gate
    // being replayed -- inside the shard; SIZE_MAX if none within the window.  Uncontrolled SWAPs further down the
    // log rename the two labels they involve.  False (and `next` untouched) outside a replay
    bool nextShardUses(const void* key, int numQubits, std::vector<size_t>& next) const {
        if (!replaying_ || window_ == 0) return false;
        next.assign((size_t) numQubits, SIZE_MAX);
        std::vector<int> nameNow((size_t) numQubits);
        for (int l = 0; l < numQubits; l++) nameNow[l] = l;
        size_t end = pos_ + 1 + window_;
        if (end > gates_.size()) end = gates_.size();
        for (size_t j = pos_ + 1; j < end; j++) {
            const LoggedGate& g = gates_[j];
            if (g.key != key) continue;
            if (g.swapA >= 0) {
                if (g.swapA < numQubits && g.swapB >= 0 && g.swapB < numQubits) std::swap(nameNow[g.swapA], nameNow[g.swapB]);
                continue;
            }
            for (int u : g.shardTargs)
                if (u >= 0 && u < numQubits && next[nameNow[u]] == SIZE_MAX) next[nameNow[u]] = j;
        }
        return true;
    }

private:
    void replay(size_t count) {
        replaying_ = true;
        for (pos_ = 0; pos_ < count; pos_++)
            gates_[pos_].run();
        gates_.erase(gates_.begin(), gates_.begin() + (std::ptrdiff_t) count);
        pos_ = 0;
        replaying_ = false;
    }

    std::deque<LoggedGate> gates_;
    size_t window_;
    size_t pos_ = 0;
    bool replaying_ = false;
};

// The victim among index bits [0, numLocalBits): not in `usedMask` (the gate's own qubits); the one whose occupant's
// next use lies farthest ahead; among equals one outside `touchedMask` (no queued gate involves it, so the backend's
// queue survives the exchange), then the least recently used.  High bits (>= 16) are tried first when the shard is
// large, so that the half-shard crossing NVLink is made of long contiguous runs.  -1 if every bit is used.
inline int chooseVictim(const std::vector<size_t>& next, const std::vector<int>& logicalAt, const std::vector<unsigned long long>& lastUse,
                        unsigned long long usedMask, unsigned long long touchedMask, int numLocalBits) {
    int victim = -1;
    const int floors[2] = {numLocalBits > 20 ? 16 : 0, 0};
    for (int lo : floors) {
        for (int p = numLocalBits - 1; p >= lo; p--) {
            if ((usedMask >> p) & 1) continue;
            if (victim < 0) { victim = p; continue; }
            size_t np = next[logicalAt[p]], nv = next[logicalAt[victim]];
            bool up = !((touchedMask >> p) & 1), uv = !((touchedMask >> victim) & 1);
            if (np != nv) { if (np > nv) victim = p; }
            else if (up != uv) { if (up) victim = p; }
            else if (lastUse[logicalAt[p]] < lastUse[logicalAt[victim]]) victim = p;
        }
        if (victim >= 0) break;
    }
    return victim;
}

}   // namespace qb_lookahead

#endif
