// b2sv developer tool (host only): prints the fusion schedule of BASELINE config 2
// (RX,RY,RZ on every wire + CNOT ring, L layers) -- passes, rounds and ops per pass.
// build: g++ -std=c++17 -O2 -I/usr/local/cuda/include -I.. schedule_dump.cpp ../schedule.cpp ../gates.cpp -o /tmp/schedule_dump
#include "schedule.hpp"
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <random>

using namespace b2sv;

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 30;
    const int layers = argc > 2 ? atoi(argv[2]) : 4;
    SchedConfig cfg;
    cfg.B = argc > 3 ? atoi(argv[3]) : 12;
    cfg.R = argc > 4 ? atoi(argv[4]) : 4;
    cfg.low = argc > 5 ? atoi(argv[5]) : 5;
    cfg.max_heavy = argc > 6 ? atoi(argv[6]) : cfg.max_heavy;
    cfg.SW = 3;
    cfg.n_local = n;
    cfg.n_alloc = n;
    std::mt19937_64 rng(42);
    std::uniform_real_distribution<double> U(0, 6.283185307179586);
    std::vector<Prim> prims;
    for (int l = 0; l < layers; l++) {
        for (int w = 0; w < n; w++)
            for (const char *g : {"RX", "RY", "RZ"})
                lower_gate(g, {n - 1 - w}, false, {U(rng)}, prims);
        for (int w = 0; w < n; w++)
            lower_gate("CNOT", {n - 1 - w, n - 1 - (w + 1) % n}, false, {}, prims);
    }
    auto passes = build_schedule(prims, cfg);
    int tot_ops = 0, tot_rounds = 0;
    for (size_t i = 0; i < passes.size(); i++) {
        const Pass &p = passes[i];
        int kinds[4] = {0, 0, 0, 0}, uncond = 0;
        for (auto &o : p.ops) {
            kinds[o.kind & OPF_KIND_MASK]++;
            uncond += (o.kind & OPF_UNCOND) ? 1 : 0;
        }
        printf("pass %2zu: fused %d ops %2d (gen %d real %d perm %d diag %d; uncond %d) rounds %d absorbed %d cx %d tile:",
               i, p.hdr.fused_store, p.hdr.n_ops, kinds[0], kinds[1], kinds[2], kinds[3], uncond, p.hdr.n_rounds,
               p.n_absorbed, p.hdr.n_cx);
        for (int j = 0; j < cfg.B; j++)
            printf(" %d", p.hdr.tile_bits[j]);
        printf("\n");
        for (int rd = 0; rd < p.hdr.n_rounds; rd++) { // bank-conflict degree of the round's gathers
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
            for (int l = 0; l < 8; l++) {
                uint32_t x = 0;
                for (int k = 0; k < 3; k++)
                    if (l & (1 << k))
                        x ^= p.hdr.round_col[rd][k] & 0xffffu;
                worst = std::max(worst, ++cnt[x & 7]);
            }
            printf("    round %d: regbits", rd);
            for (int s = 0; s < cfg.R; s++)
                printf(" %d", p.hdr.round_regbits[rd][s]);
            printf("  kind %d  ops %d  gather conflict degree %d\n", p.hdr.round_kind[rd], p.hdr.round_begin[rd + 1] - p.hdr.round_begin[rd], worst);
        }
        tot_ops += p.hdr.n_ops;
        tot_rounds += p.hdr.n_rounds;
    }
    printf("%zu passes, %d ops, %d rounds\n", passes.size(), tot_ops, tot_rounds);
    return 0;
}
