// b2sv TEST INFRASTRUCTURE (host only, never linked into libb2sv.so): a scalar re-execution of the
// pass descriptors the fusion scheduler emits, following the tile executor (tile_kernel.cu) step by
// step -- tile-id deposit, swizzled tile slots, per-round gather through the GF(2)-affine address
// map, the op interpreter, dense and factored rounds, conditional address toggles, the fused store
// and the store phase. It lets the CPU test suite check that  lower -> schedule -> descriptors
// reproduces the oracle's amplitudes for every gate and scheduling decision without a GPU; it is
// not a fallback (the product library has no reference to it).
// build: see tests/conftest.py (g++ -shared with ../schedule.cpp ../gates.cpp)
#include "schedule.hpp"

#include <cstdlib>
#include <cstring>
#include <string>

using namespace b2sv;

namespace {

template <typename real> struct Amp {
    real x, y;
};

template <typename real> struct Emu {
    using amp = Amp<real>;
    int B, R, SW, SH, low, n_eff;
    std::vector<amp> &state;

    uint32_t phys(uint32_t i) const { return phys_slot(i, B, SW, SH); }

    void run_op(amp *a, const DevOp &op, uint64_t tbr, uint32_t base) const {
        const int NS = 1 << R;
        real m[8];
        for (int i = 0; i < 8; i++)
            m[i] = static_cast<real>(op.m[i]);
        const int kind = op.kind & OPF_KIND_MASK;
        const int ts = op.tslot;
        auto general = [&](int s, int s1) {
            const amp v0 = a[s], v1 = a[s1];
            a[s].x = m[0] * v0.x - m[1] * v0.y + m[2] * v1.x - m[3] * v1.y;
            a[s].y = m[0] * v0.y + m[1] * v0.x + m[2] * v1.y + m[3] * v1.x;
            a[s1].x = m[4] * v0.x - m[5] * v0.y + m[6] * v1.x - m[7] * v1.y;
            a[s1].y = m[4] * v0.y + m[5] * v0.x + m[6] * v1.y + m[7] * v1.x;
        };
        if ((op.kind & OPF_UNCOND) && kind <= KIND_REAL) {
            for (int s = 0; s < NS; s++)
                if (!(s & (1 << ts)))
                    general(s, s | (1 << ts));
            return;
        }
        if ((tbr & op.gcm) != op.gcv)
            return;
        const bool pred = (base & op.lcm) == op.lcv;
        if (!pred)
            return;
        if (kind == KIND_DIAG) {
            const bool odd0 = ((__builtin_popcountll(tbr & op.gpm) + __builtin_popcount(base & op.lpm)) & 1) != 0;
            for (int s = 0; s < NS; s++) {
                if (!((op.slot_act >> s) & 1u))
                    continue;
                const bool odd = odd0 ^ (((op.slot_par >> s) & 1u) != 0);
                const real pr = odd ? m[2] : m[0], pi = odd ? m[3] : m[1];
                const amp v = a[s];
                a[s].x = pr * v.x - pi * v.y;
                a[s].y = pr * v.y + pi * v.x;
            }
            return;
        }
        for (int s = 0; s < NS; s++) {
            if ((s & (1 << ts)) || !((op.slot_act >> s) & 1u))
                continue;
            const int s1 = s | (1 << ts);
            if (kind == KIND_PERM)
                std::swap(a[s], a[s1]);
            else
                general(s, s1);
        }
    }

    void dense_factored(amp *a, int G, const DevDense &dd) const {
        const int NS = 1 << R, GM = (1 << G) - 1;
        const amp *pre = reinterpret_cast<const amp *>(dd.pre);
        const amp *post = reinterpret_cast<const amp *>(dd.post);
        const real *t = reinterpret_cast<const real *>(dd.t);
        for (int s = 0; s < NS; s++) {
            if ((s & GM) == 0)
                continue;
            const amp p = pre[s & GM], v = a[s];
            a[s].x = p.x * v.x - p.y * v.y;
            a[s].y = p.x * v.y + p.y * v.x;
        }
        for (int k = 0; k < G; k++)
            for (int s = 0; s < NS; s++) {
                if (s & (1 << k))
                    continue;
                const int s1 = s | (1 << k);
                const amp v0 = a[s], v1 = a[s1];
                a[s].x = v0.x + t[2 * k] * v1.x;
                a[s].y = v0.y + t[2 * k] * v1.y;
                a[s1].x = v1.x + t[2 * k + 1] * v0.x;
                a[s1].y = v1.y + t[2 * k + 1] * v0.y;
            }
        for (int s = 0; s < NS; s++) {
            const amp p = post[s & GM], v = a[s];
            a[s].x = p.x * v.x - p.y * v.y;
            a[s].y = p.x * v.y + p.y * v.x;
        }
    }

    void run_pass(const Pass &ps, uint64_t rank_bits) {
        const DevPassHeader &h = ps.hdr;
        const int NS = 1 << R, NF = B - R, GT = 1 << NF, TILE = 1 << B;
        const uint32_t n_tiles = 1u << (n_eff - B);
        const uint32_t lowmask = (1u << h.low_bits) - 1u;
        std::vector<uint64_t> rowoff(size_t(1) << (B - h.low_bits));
        for (size_t r = 0; r < rowoff.size(); r++) {
            uint64_t off = 0;
            for (int j = h.low_bits; j < B; j++)
                if ((r >> (j - h.low_bits)) & 1)
                    off |= uint64_t(1) << h.tile_bits[j];
            rowoff[r] = off;
        }
        std::vector<amp> tile(TILE), a(NS);
        for (uint32_t t = 0; t < n_tiles; t++) {
            uint64_t tb = 0;
            for (int j = 0; j < h.n_seg; j++)
                tb |= static_cast<uint64_t>(t & h.seg_mask[j]) << h.seg_shift[j];
            const uint64_t tbr = tb | rank_bits;
            for (uint32_t i = 0; i < static_cast<uint32_t>(TILE); i++)
                tile[h.plain_layout ? i : phys(i)] = state[tb | rowoff[i >> h.low_bits] | (i & lowmask)];
            std::vector<uint32_t> xoff(h.n_rounds + 1, 0);
            for (int r = 0; r <= h.n_rounds; r++)
                for (int c = 0; c < h.n_cx; c++)
                    if (h.cx[c].round <= r && (tbr & h.cx[c].gcm) == h.cx[c].gcv)
                        xoff[r] ^= h.cx[c].vec;
            for (int rd = 0; rd < h.n_rounds; rd++) {
                const int fused = rd == h.n_rounds - 1 ? h.fused_store : 0;
                const int kind = h.round_kind[rd];
                const int ob = h.round_begin[rd], oe = h.round_begin[rd + 1];
                for (int tid = 0; tid < GT; tid++) {
                    uint32_t acc = 0;
                    for (int c = 0; c < NF; c++)
                        if ((tid >> c) & 1)
                            acc ^= h.round_col[rd][c];
                    const uint32_t base = acc >> 16;
                    const uint32_t pb = (acc & 0xffffu) ^ xoff[rd];
                    auto slot_addr = [&](int s) {
                        uint32_t x = pb;
                        for (int c = 0; c < R; c++)
                            if (s & (1 << c))
                                x ^= h.round_poff[rd][c];
                        return x;
                    };
                    for (int s = 0; s < NS; s++)
                        a[s] = tile[slot_addr(s)];
                    if (kind == 0) {
                        for (int oi = ob; oi < oe; oi++)
                            run_op(a.data(), ps.ops[oi], tbr, base);
                    } else if (kind >= 8) {
                        dense_factored(a.data(), kind - 8, ps.dense[rd]);
                    } else {
                        for (int k = 0; k < kind; k++) {
                            DevOp op = ps.ops[ob + k];
                            B2_ASSERT((op.kind & OPF_UNCOND) && op.tslot == k);
                            run_op(a.data(), op, tbr, base);
                        }
                    }
                    if (fused == 1) {
                        uint64_t addr0 = tb;
                        for (int c = 0; c < NF; c++)
                            if ((tid >> c) & 1)
                                addr0 ^= h.store_free[c];
                        for (int c = 0; c < h.n_cx; c++)
                            if ((tbr & h.cx[c].gcm) == h.cx[c].gcv)
                                addr0 ^= h.store_cx[c];
                        for (int s = 0; s < NS; s++) {
                            uint64_t x = addr0;
                            for (int c = 0; c < R; c++)
                                if (s & (1 << c))
                                    x ^= h.store_reg[c];
                            state[x] = a[s];
                        }
                    } else {
                        for (int s = 0; s < NS; s++)
                            tile[slot_addr(s)] = a[s];
                    }
                }
            }
            if (h.fused_store && h.n_rounds > 0)
                continue;
            for (uint32_t i = 0; i < static_cast<uint32_t>(TILE); i++) {
                uint32_t x = xoff[h.n_rounds];
                for (int j = 0; j < B; j++)
                    if ((i >> j) & 1)
                        x ^= h.final_col[j];
                state[tb | rowoff[i >> h.low_bits] | (i & lowmask)] = tile[x];
            }
        }
    }
};

template <typename real>
int run(int n, int B, int R, int low, int max_heavy, int factor, int store_mode,
        const std::vector<Prim> &prims, double *st, uint64_t *stats) {
    SchedConfig cfg;
    cfg.B = B;
    cfg.R = R;
    cfg.low = low;
    cfg.SW = sizeof(real) == 8 ? 3 : 4;
    cfg.SH = sizeof(real) == 8 ? 0 : 1;
    cfg.f32 = sizeof(real) == 4;
    cfg.factor = factor != 0;
    cfg.max_heavy = max_heavy;
    cfg.lookahead = getenv("B2EMU_NO_LOOKAHEAD") == nullptr;
    cfg.bulk = getenv("B2EMU_BULK") != nullptr;
    cfg.bulk_min_run_bits = low; // the tests want as many plain-layout passes as possible
    cfg.fuse_store = store_mode != 0;
    cfg.n_local = n;
    cfg.n_alloc = std::max(n, B);
    const int n_eff = cfg.n_alloc;
    std::vector<Amp<real>> state(size_t(1) << n_eff, Amp<real>{0, 0});
    for (size_t i = 0; i < (size_t(1) << n); i++)
        state[i] = {static_cast<real>(st[2 * i]), static_cast<real>(st[2 * i + 1])};
    Emu<real> emu{B, R, cfg.SW, cfg.SH, low, n_eff, state};
    uint64_t n_pass = 0, n_fact = 0, n_dense = 0, n_rounds = 0, n_fused = 0, n_staged = 0;
    for (const Pass &ps : build_schedule(prims, cfg)) {
        B2_ABORT_IF(ps.is_matk, "emulator: generic k-qubit matrices are not covered");
        emu.run_pass(ps, 0);
        n_pass++;
        n_fused += ps.hdr.fused_store == 1;
        n_staged += ps.hdr.plain_layout == 1; // (slot 5 of the stats: passes in the plain tile layout)
        for (int rd = 0; rd < ps.hdr.n_rounds; rd++) {
            n_rounds++;
            n_fact += ps.hdr.round_kind[rd] >= 8;
            n_dense += ps.hdr.round_kind[rd] >= 1 && ps.hdr.round_kind[rd] < 8;
        }
    }
    for (size_t i = 0; i < (size_t(1) << n); i++) {
        st[2 * i] = state[i].x;
        st[2 * i + 1] = state[i].y;
    }
    if (stats) {
        stats[0] = n_pass;
        stats[1] = n_rounds;
        stats[2] = n_dense;
        stats[3] = n_fact;
        stats[4] = n_fused;
        stats[5] = n_staged;
    }
    return 0;
}

thread_local std::string g_err;

} // namespace

extern "C" {

const char *b2emu_last_error() { return g_err.c_str(); }

// the shared-memory swizzle of schedule.hpp, for the property tests
uint32_t b2emu_phys_slot(uint32_t i, int B, int SW, int SH) { return phys_slot(i, B, SW, SH); }

// names / wires / params as in b2sv_ops_create; state: 2^n interleaved (re, im) doubles, in place.
// stats (6 values): passes, rounds, dense rounds, factored rounds, fused stores, plain-layout passes.
// store_mode: 0 = always the store phase, 1 = fused stores where the schedule allows.
int b2emu_run(int n, int f32, int B, int R, int low, int max_heavy, int factor, int store_mode,
              int n_ops,
              const char **names, const int64_t *wires_flat, const int *nw, const int *inverse,
              const double *params_flat, const int *np, double *state, uint64_t *stats) {
    try {
        std::vector<Prim> prims;
        size_t wo = 0, po = 0;
        for (int i = 0; i < n_ops; i++) {
            std::vector<int64_t> w(wires_flat + wo, wires_flat + wo + nw[i]);
            std::vector<double> p(params_flat + po, params_flat + po + np[i]);
            wo += nw[i];
            po += np[i];
            if (std::string(names[i]) == "Identity")
                continue;
            const bool ok = lower_gate(names[i], wires_to_bits(w, n), inverse[i] != 0, p, prims);
            B2_ABORT_IF(!ok, std::string("emulator: not a named gate: ") + names[i]);
        }
        if (prims.empty())
            return 0;
        return f32 ? run<float>(n, B, R, low, max_heavy, factor, store_mode, prims, state, stats)
                   : run<double>(n, B, R, low, max_heavy, factor, store_mode, prims, state, stats);
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
}
