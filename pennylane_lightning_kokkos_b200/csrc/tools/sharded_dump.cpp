// b2sv developer tool (host only): replays State::apply_prims_sharded's segmentation for BASELINE
// config 2 on 2^g ranks and prints, per segment, the ops, passes and the swaps chosen.
// build: g++ -std=c++17 -O2 -I/usr/local/cuda/include -I.. sharded_dump.cpp ../schedule.cpp ../gates.cpp -o /tmp/sharded_dump
#include "schedule.hpp"
#include <cstdio>
#include <cstdlib>
#include <random>
using namespace b2sv;
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 31, g = argc > 2 ? atoi(argv[2]) : 1, layers = argc > 3 ? atoi(argv[3]) : 4;
    const int n_local = n - g;
    std::mt19937_64 rng(42);
    std::uniform_real_distribution<double> U(0, 6.283185307179586);
    std::vector<Prim> pending;
    for (int l = 0; l < layers; l++) {
        for (int w = 0; w < n; w++)
            for (const char *gname : {"RX", "RY", "RZ"})
                lower_gate(gname, {n - 1 - w}, false, {U(rng)}, pending);
        for (int w = 0; w < n; w++)
            lower_gate("CNOT", {n - 1 - w, n - 1 - (w + 1) % n}, false, {}, pending);
    }
    std::vector<int> l2p(n);
    for (int q = 0; q < n; q++) l2p[q] = q;
    auto phys_mask = [&](uint64_t m) { uint64_t r = 0; while (m) { int q = __builtin_ctzll(m); m &= m - 1; r |= bit(l2p[q]); } return r; };
    int tot_pass = 0, tot_swaps = 0, it = 0;
    while (!pending.empty()) {
        std::vector<Prim> seg, rest;
        uint64_t T = 0, D = 0;
        for (const Prim &p : pending) {
            const uint64_t tm = p.target_mask(), dm = p.support() & ~tm;
            const bool blocked = (tm & (T | D)) || (dm & T);
            const bool local = (phys_mask(tm) >> n_local) == 0;
            if (!blocked && local) {
                Prim q = p;
                if (p.type == Prim::C1Q) q.target = l2p[p.target];
                q.cmask = phys_mask(p.cmask); q.cval = phys_mask(p.cval); q.pmask = phys_mask(p.pmask);
                seg.push_back(q);
            } else { T |= tm; D |= dm; rest.push_back(p); }
        }
        int np = 0, nr = 0;
        if (!seg.empty()) {
            SchedConfig cfg; cfg.B = 12; cfg.R = 4; cfg.SW = 3; cfg.low = 5; cfg.n_local = n_local; cfg.n_alloc = n_local;
            for (auto &ps : build_schedule(seg, cfg)) { np++; nr += ps.hdr.n_rounds; }
        }
        tot_pass += np;
        printf("iter %d: segment %zu prims -> %d passes %d rounds; %zu pending", it++, seg.size(), np, nr, rest.size());
        if (rest.empty()) { printf("\n"); break; }
        std::vector<int> need; uint64_t need_mask = 0;
        for (const Prim &p : rest) {
            uint64_t tm = p.target_mask();
            while (tm && (int)need.size() < g) { int q = __builtin_ctzll(tm); tm &= tm - 1; if (l2p[q] >= n_local && !((need_mask >> q) & 1)) { need.push_back(q); need_mask |= bit(q); } }
            if ((int)need.size() >= g) break;
        }
        std::vector<size_t> next_use(n, rest.size() + 1);
        for (size_t i = rest.size(); i-- > 0;) { uint64_t tm = rest[i].target_mask(); while (tm) { next_use[__builtin_ctzll(tm)] = i; tm &= tm - 1; } }
        for (int q : need) {
            int victim = -1;
            for (int min_pos : {5, 0}) {
            for (int o = 0; o < n; o++) {
                if (l2p[o] >= n_local || l2p[o] < min_pos || ((need_mask >> o) & 1)) continue;
                if (victim < 0 || next_use[o] > next_use[victim] || (next_use[o] == next_use[victim] && l2p[o] > l2p[victim])) victim = o;
            }
            if (victim >= 0) break; }
            printf("; swap logical %d (global) <-> logical %d (phys %d, next use %zu)", q, victim, l2p[victim], next_use[victim]);
            std::swap(l2p[q], l2p[victim]);
            tot_swaps++;
        }
        printf("\n");
        pending.swap(rest);
    }
    printf("total: %d passes, %d swaps\n", tot_pass, tot_swaps);
}
