// CPU harness for quest_b200/shim/lookahead.hpp (tests/test_lookahead_cpu.py builds and drives it).
// Reads "<numQubits> <numLocalBits> <window>" then one gate per line:
//   g <k> <q1> .. <qk> <m> <o1> .. <om>   k non-diagonal targets, m other qubits the gate touches
//   s <a> <b>                             uncontrolled SWAP (relabelling only)
//   r                                     something else touches the state: drain the log
// and keeps the same bookkeeping as the shim's qubit map (localiser_b200.cpp: QubitMap, touch, pullTargetsIntoShard)
// with the exchange replaced by a counter.  Prints "<exchanges> <gates run> <max log size> <order ok>".
#include "lookahead.hpp"

#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

struct Map {
    std::vector<int> phys, logi;
    std::vector<unsigned long long> lastUse;
    unsigned long long clock = 0;
};

int main() {
    int n, nl, window;
    if (!(std::cin >> n >> nl >> window)) return 2;
    Map m;
    m.phys.resize(n); m.logi.resize(n); m.lastUse.assign(n, 0);
    for (int i = 0; i < n; i++) m.phys[i] = m.logi[i] = i;
    qb_lookahead::GateLog log(window);
    long exchanges = 0, ran = 0, issued = 0, maxLog = 0;
    bool orderOk = true;
    const int key = 0;

    auto runGate = [&](std::vector<int> nd, std::vector<int> other, long id) {
        if (id != ran) orderOk = false;
        ran++;
        unsigned long long used = 0;
        for (int& q : other) { m.lastUse[q] = ++m.clock; q = m.phys[q]; used |= 1ULL << q; }
        for (int& q : nd) { m.lastUse[q] = ++m.clock; q = m.phys[q]; used |= 1ULL << q; }
        for (int& t : nd) {
            if (t < nl) continue;
            int victim = -1;
            std::vector<size_t> next;
            if (log.nextShardUses(&key, n, next))
                victim = qb_lookahead::chooseVictim(next, m.logi, m.lastUse, used, 0ULL, nl);
            else {                                            // the shim's default: least recently used, high bits first
                for (int lo : {nl > 20 ? 16 : 0, 0}) {
                    for (int p = nl - 1; p >= lo; p--) {
                        if ((used >> p) & 1) continue;
                        if (victim < 0 || m.lastUse[m.logi[p]] < m.lastUse[m.logi[victim]]) victim = p;
                    }
                    if (victim >= 0) break;
                }
            }
            if (victim < 0) { std::puts("no victim"); std::exit(3); }
            int lt = m.logi[t], lv = m.logi[victim];
            m.logi[t] = lv; m.logi[victim] = lt;
            m.phys[lt] = victim; m.phys[lv] = t;
            used = (used & ~(1ULL << t)) | (1ULL << victim);
            t = victim;
            exchanges++;
        }
    };
    auto runSwap = [&](int a, int b, long id) {
        if (id != ran) orderOk = false;
        ran++;
        int pa = m.phys[a], pb = m.phys[b];
        m.phys[a] = pb; m.phys[b] = pa; m.logi[pa] = b; m.logi[pb] = a;
    };

    std::string line;
    std::getline(std::cin, line);
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::istringstream in(line);
        char kind; in >> kind;
        if (kind == 'r') { log.drain(); if (!log.empty()) orderOk = false; continue; }
        long id = issued++;
        if (kind == 's') {
            int a, b; in >> a >> b;
            if (window > 0) log.push(&key, {}, [=]() { runSwap(a, b, id); }, a, b); else runSwap(a, b, id);
        } else {
            int k, mm; in >> k; std::vector<int> nd(k); for (int& q : nd) in >> q;
            in >> mm; std::vector<int> other(mm); for (int& q : other) in >> q;
            if (window > 0) log.push(&key, nd, [=]() { runGate(nd, other, id); }); else runGate(nd, other, id);
        }
        if ((long) log.size() > maxLog) maxLog = (long) log.size();
    }
    log.drain();
    std::printf("%ld %ld %ld %d\n", exchanges, ran, maxLog, (orderOk && ran == issued && log.empty()) ? 1 : 0);
    return 0;
}
