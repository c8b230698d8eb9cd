// b2sv: fusion scheduler implementation (see schedule.hpp).
#include "schedule.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b2sv {
namespace {

inline int popc(uint64_t x) { return __builtin_popcountll(x); }
inline int ctz(uint64_t x) { return __builtin_ctzll(x); }

bool is_plain_1q(const Prim &p) {
    if (p.tag >= 0 || p.cmask != 0)
        return false;
    if (p.type == Prim::C1Q)
        return true;
    return p.type == Prim::DIAG && popc(p.pmask) == 1;
}
int bit_of_1q(const Prim &p) { return p.type == Prim::C1Q ? p.target : ctz(p.pmask); }
void matrix_of_1q(const Prim &p, cplx m[4]) {
    if (p.type == Prim::C1Q) {
        for (int i = 0; i < 4; i++)
            m[i] = p.m[i];
    } else {
        m[0] = p.m[0];
        m[1] = 0;
        m[2] = 0;
        m[3] = p.m[1];
    }
}
void store_1q(Prim &p, int t, const cplx m[4]) {
    p.cmask = p.cval = 0;
    if (m[1] == cplx(0.0) && m[2] == cplx(0.0)) {
        p.type = Prim::DIAG;
        p.target = -1;
        p.pmask = bit(t);
        p.m[0] = m[0];
        p.m[1] = m[3];
        p.m[2] = p.m[3] = 0;
    } else {
        p.type = Prim::C1Q;
        p.target = t;
        p.pmask = 0;
        for (int i = 0; i < 4; i++)
            p.m[i] = m[i];
    }
}
} // namespace

std::vector<Prim> fuse_single_qubit(const std::vector<Prim> &prims) {
    std::vector<Prim> out;
    out.reserve(prims.size());
    int last_touch[64];
    for (int &x : last_touch)
        x = -1;
    for (const Prim &p : prims) {
        if (is_plain_1q(p)) {
            const int t = bit_of_1q(p);
            const int j = last_touch[t];
            if (j >= 0 && is_plain_1q(out[j]) && bit_of_1q(out[j]) == t) {
                // nothing between out[j] and p touches bit t, so p commutes back to out[j]
                cplx a[4], b[4], c[4];
                matrix_of_1q(p, a);
                matrix_of_1q(out[j], b);
                c[0] = a[0] * b[0] + a[1] * b[2];
                c[1] = a[0] * b[1] + a[1] * b[3];
                c[2] = a[2] * b[0] + a[3] * b[2];
                c[3] = a[2] * b[1] + a[3] * b[3];
                store_1q(out[j], t, c);
                continue;
            }
        }
        const int idx = static_cast<int>(out.size());
        out.push_back(p);
        if (out.back().type == Prim::C1Q && out.back().m[1] == cplx(0.0) &&
            out.back().m[2] == cplx(0.0)) { // diagonal 2x2 -> DIAG (keeps its control)
            Prim &q = out.back();
            q.type = Prim::DIAG;
            q.pmask = bit(q.target);
            q.target = -1;
            q.m[1] = q.m[3];
            q.m[2] = q.m[3] = 0;
        }
        uint64_t s = out.back().support();
        while (s) {
            last_touch[ctz(s)] = idx;
            s &= s - 1;
        }
    }
    return out;
}

namespace {

OpKind classify(const Prim &p) {
    if (p.type == Prim::DIAG)
        return KIND_DIAG;
    const cplx *m = p.m;
    if (m[0] == cplx(0.0) && m[3] == cplx(0.0) && m[1] == cplx(1.0) && m[2] == cplx(1.0))
        return KIND_PERM;
    if (m[0].imag() == 0 && m[1].imag() == 0 && m[2].imag() == 0 && m[3].imag() == 0)
        return KIND_REAL;
    return KIND_GENERAL;
}

// ---- factored form of a 2x2 (see DevDense) ---------------------------------------------------------
// M = diag(P0, P1) * [[1, t01], [t10, 1]] * diag(1, q1) with |q1| = 1 and t01, t10 real:
//   out0 = P0 * (v0 + t01 * (q1 v1)),   out1 = P1 * ((q1 v1) + t10 * v0).
// Exists iff m00, m11 != 0 and arg(m01 m10) = arg(m00 m11) (mod pi) -- true for every unitary --
// and is accepted when both shears are moderate (|t| <= kShearMax; unitaries with a dominant
// diagonal have |t| <= 1) and the neglected imaginary part is at rounding level.
struct Factored {
    cplx q1, P0, P1;
    double t01, t10;
};
constexpr double kShearMax = 2.0;
constexpr double kShearImagTol = 32.0 * 2.220446049250313e-16;

bool factor_2x2(const cplx m[4], Factored &f) {
    const cplx a = m[0], b = m[1], c = m[2], d = m[3];
    const double na = std::abs(a), nd = std::abs(d), nb = std::abs(b), nc = std::abs(c);
    if (!(na > 0.0) || !(nd > 0.0) || !std::isfinite(na + nd + nb + nc))
        return false;
    if (nb > kShearMax * na || nc > kShearMax * nd)
        return false;
    cplx q1(1.0, 0.0);
    if (nb >= nc && nb > 0.0) {
        const cplx z = b / a;
        q1 = z / std::abs(z);
    } else if (nc > 0.0) {
        const cplx z = c / d;
        q1 = std::conj(z) / std::abs(z);
    }
    const cplx t01 = b / (a * q1), t10 = c * q1 / d;
    if (std::abs(t01.imag()) > kShearImagTol * (1.0 + std::abs(t01)) ||
        std::abs(t10.imag()) > kShearImagTol * (1.0 + std::abs(t10)))
        return false;
    f.q1 = q1;
    f.P0 = a;
    f.P1 = d / q1;
    f.t01 = t01.real();
    f.t10 = t10.real();
    return true;
}

bool is_free_2x2(const Prim &p) { return p.type == Prim::C1Q && p.cmask == 0 && p.tag < 0; }

// An uncontrolled 2x2 whose anti-diagonal dominates is rewritten as X * (X M): X M has the dominant
// diagonal the factored form wants, and the X is absorbed into the tile's address map for free.
std::vector<Prim> split_antidiagonal(const std::vector<Prim> &prims) {
    std::vector<Prim> out;
    out.reserve(prims.size() + prims.size() / 2);
    for (const Prim &p : prims) {
        if (!is_free_2x2(p) || classify(p) == KIND_PERM ||
            std::abs(p.m[1]) * std::abs(p.m[2]) <= std::abs(p.m[0]) * std::abs(p.m[3])) {
            out.push_back(p);
            continue;
        }
        const cplx xm[4] = {p.m[2], p.m[3], p.m[0], p.m[1]}; // X M: rows swapped
        Factored f;
        if (!factor_2x2(xm, f)) {
            out.push_back(p);
            continue;
        }
        Prim q = p;
        if (xm[1] == cplx(0.0) && xm[2] == cplx(0.0)) { // X M is diagonal (e.g. PauliY)
            q.type = Prim::DIAG;
            q.pmask = bit(p.target);
            q.target = -1;
            q.m[0] = xm[0];
            q.m[1] = xm[3];
            q.m[2] = q.m[3] = 0;
        } else {
            for (int i = 0; i < 4; i++)
                q.m[i] = xm[i];
        }
        out.push_back(q);
        Prim x;
        x.type = Prim::C1Q;
        x.target = p.target;
        x.m[0] = x.m[3] = 0;
        x.m[1] = x.m[2] = 1;
        out.push_back(x);
    }
    return out;
}

template <typename real>
void fill_dense(DevDense &dd, const std::vector<Factored> &fs) {
    std::memset(&dd, 0, sizeof(dd));
    const int G = static_cast<int>(fs.size());
    real *pre = reinterpret_cast<real *>(dd.pre), *post = reinterpret_cast<real *>(dd.post);
    real *t = reinterpret_cast<real *>(dd.t);
    for (int j = 0; j < (1 << G); j++) {
        cplx a(1.0, 0.0), b(1.0, 0.0);
        for (int k = 0; k < G; k++) {
            if ((j >> k) & 1) {
                a *= fs[k].q1;
                b *= fs[k].P1;
            } else {
                b *= fs[k].P0;
            }
        }
        pre[2 * j] = static_cast<real>(a.real());
        pre[2 * j + 1] = static_cast<real>(a.imag());
        post[2 * j] = static_cast<real>(b.real());
        post[2 * j + 1] = static_cast<real>(b.imag());
    }
    for (int k = 0; k < G; k++) {
        t[2 * k] = static_cast<real>(fs[k].t01);
        t[2 * k + 1] = static_cast<real>(fs[k].t10);
    }
}

// Dependency filter shared by pass- and round-level greedy selection.
struct Blocker {
    uint64_t t = 0; // bits some skipped op acts on non-diagonally
    uint64_t d = 0; // bits some skipped op uses diagonally (control / phase)
    bool blocked(const Prim &p) const {
        const uint64_t tm = p.target_mask(), dm = p.support() & ~tm;
        return (tm & (t | d)) || (dm & t);
    }
    void skip(const Prim &p) {
        const uint64_t tm = p.target_mask();
        t |= tm;
        d |= p.support() & ~tm;
    }
};

DevOp make_devop(const Prim &p, const uint8_t *tile_bits, int B, const uint8_t *regbits, int R,
                 int jac) {
    DevOp o;
    std::memset(&o, 0, sizeof(o));
    o.kind = classify(p);
    o.jac = static_cast<int16_t>(jac);
    if (p.type == Prim::DIAG) {
        o.m[0] = p.m[0].real();
        o.m[1] = p.m[0].imag();
        o.m[2] = p.m[1].real();
        o.m[3] = p.m[1].imag();
    } else {
        for (int i = 0; i < 4; i++) {
            o.m[2 * i] = p.m[i].real();
            o.m[2 * i + 1] = p.m[i].imag();
        }
    }
    uint64_t tile_mask = 0;
    for (int j = 0; j < B; j++)
        tile_mask |= bit(tile_bits[j]);
    o.gcm = p.cmask & ~tile_mask;
    o.gcv = p.cval & ~tile_mask;
    o.gpm = p.pmask & ~tile_mask;
    uint32_t rcm = 0, rcv = 0, rpm = 0;
    int tslot = -1;
    for (int j = 0; j < B; j++) {
        const uint64_t g = bit(tile_bits[j]);
        int slot = -1;
        for (int s = 0; s < R; s++)
            if (regbits[s] == j)
                slot = s;
        if (p.type == Prim::C1Q && p.target == tile_bits[j]) {
            B2_ASSERT(slot >= 0);
            tslot = slot;
        }
        if (p.cmask & g) {
            if (slot >= 0) {
                rcm |= 1u << slot;
                if (p.cval & g)
                    rcv |= 1u << slot;
            } else {
                o.lcm |= 1u << j;
                if (p.cval & g)
                    o.lcv |= 1u << j;
            }
        }
        if (p.pmask & g) {
            if (slot >= 0)
                rpm |= 1u << slot;
            else
                o.lpm |= 1u << j;
        }
    }
    if (p.type == Prim::C1Q) {
        B2_ASSERT(tslot >= 0);
        o.tslot = static_cast<uint8_t>(tslot);
    }
    for (uint32_t s = 0; s < (1u << R); s++) {
        if ((s & rcm) == rcv)
            o.slot_act |= 1u << s;
        if (__builtin_popcount(s & rpm) & 1)
            o.slot_par |= 1u << s;
    }
    if (p.cmask == 0)
        o.kind |= OPF_UNCOND;
    return o;
}

} // namespace

// tile-id deposit segments: one per run of consecutive index bits below `top` that are not excluded
void fill_tile_id_segments(DevPassHeader &hdr, uint64_t excluded_mask, int top) {
    int n_seg = 0, id_bit = 0, below = 0; // below = excluded bits under the current position
    int b = 0;
    while (b < top) {
        if (excluded_mask & bit(b)) {
            below++;
            b++;
            continue;
        }
        int len = 0;
        while (b + len < top && !(excluded_mask & bit(b + len)))
            len++;
        B2_ASSERT(n_seg <= kMaxTileBits);
        hdr.seg_mask[n_seg] = static_cast<uint32_t>(((uint64_t(1) << len) - 1) << id_bit);
        hdr.seg_shift[n_seg] = static_cast<uint8_t>(below);
        n_seg++;
        id_bit += len;
        b += len;
    }
    for (int k = n_seg; k <= kMaxTileBits; k++) {
        hdr.seg_mask[k] = 0;
        hdr.seg_shift[k] = 0;
    }
    hdr.n_seg = static_cast<uint8_t>(n_seg);
}

int default_tile_low() {
    const char *e = getenv("B2SV_TILE_LOW"); // read on every call: experiments sweep it
    return e ? std::max(3, std::min(6, atoi(e))) : kDefaultTileLow; // 3 = the tile kernel's kMinLow
}

double schedule_cost(const std::vector<Pass> &passes) {
    // measured (profiles/r1_tile_ablation.md): 1 round 6.3-6.6 ms, +0.8-1.0 ms per further round
    double c = 0.0;
    for (const Pass &ps : passes)
        c += ps.is_matk ? 8.0 : 5.6 + 0.9 * std::max(1, static_cast<int>(ps.hdr.n_rounds));
    return c;
}

namespace {
std::vector<Pass> build_schedule_fixed(const std::vector<Prim> &prims_in, const SchedConfig &cfg);
constexpr int kAutoBudgetMinBits = 25; // from 2^25 amplitudes on, trying several budgets pays
}

std::vector<Pass> build_schedule(const std::vector<Prim> &prims_in, const SchedConfig &cfg_in) {
    SchedConfig cfg = cfg_in;
    if (const char *e = getenv("B2SV_LOOKAHEAD")) // experiments: tile bits by marginal gain
        cfg.lookahead = atoi(e) != 0;
    if (cfg.max_heavy > 0)
        return build_schedule_fixed(prims_in, cfg);
    if (std::max(cfg.n_alloc, cfg.n_local) < kAutoBudgetMinBits) {
        // small states: a sweep costs microseconds, so host time and launch count matter more than
        // the balance inside a pass -- one cheap greedy build with a generous budget
        SchedConfig c = cfg;
        c.max_heavy = 16;
        return build_schedule_fixed(prims_in, c);
    }
    std::vector<Pass> best;
    double best_cost = 0.0;
    for (int mh : {8, 10, 12, 16, 24}) {
        SchedConfig c = cfg;
        c.max_heavy = mh;
        std::vector<Pass> s = build_schedule_fixed(prims_in, c);
        const double cost = schedule_cost(s);
        if (best.empty() || cost < best_cost) {
            best = std::move(s);
            best_cost = cost;
        }
        if (static_cast<int>(prims_in.size()) <= mh)
            break; // larger budgets cannot change anything
    }
    return best;
}

namespace {
std::vector<Pass> build_schedule_fixed(const std::vector<Prim> &prims_in, const SchedConfig &cfg) {
    const int B = cfg.B, R = cfg.R, low = std::min(cfg.low, cfg.B);
    B2_ASSERT(B <= kMaxTileBits && R <= kMaxRegBits && R <= B);
    const bool factor = cfg.factor && cfg.free_perms && cfg.fuse;
    const std::vector<Prim> prims =
        factor ? split_antidiagonal(fuse_single_qubit(prims_in)) : fuse_single_qubit(prims_in);
    const int N = static_cast<int>(prims.size());
    std::vector<char> done(N, 0);
    // permutation 2x2s (X, CNOT, Toffoli ...) that may be folded into the address map: classified once
    std::vector<char> is_perm(N, 0);
    for (int i = 0; i < N; i++)
        is_perm[i] = prims[i].type == Prim::C1Q && prims[i].tag < 0 && classify(prims[i]) == KIND_PERM;
    std::vector<Pass> passes;
    passes.reserve(static_cast<size_t>(N) / 4 + 8); // a Pass carries a 4 KB header: avoid regrowth copies
    const uint64_t low_mask = bit(low) - 1;

    int first = 0;
    while (first < N) {
        if (done[first]) {
            first++;
            continue;
        }
        if (prims[first].type == Prim::MATK) {
            Pass ps;
            ps.is_matk = true;
            ps.matk = prims[first];
            passes.push_back(std::move(ps));
            done[first++] = 1;
            continue;
        }
        // ---- choose the tile bits of this pass, then its ops
        // select(T, window, out): scan the pending ops in program order; an op joins the pass when
        // nothing it depends on was skipped (Blocker), its target is a tile bit and the arithmetic
        // budget allows. Returns 64 * (arithmetic ops) + (free permutations).
        // With grow > 0 the tile takes up to `grow` further target bits first come, first served.
        auto select = [&](uint64_t tmask, int grow, int window, std::vector<int> *out,
                          uint64_t *tmask_out) {
            Blocker blk;
            int heavy = 0, n_light = 0, n_ops = 0, seen = 0, misses = 0;
            // after kMaxMisses pending ops in a row that could not join, the pass is taken as full
            constexpr int kMaxMisses = 512;
            for (int i = first; i < N && n_ops < kMaxOpsPerPass && seen < window && misses < kMaxMisses; i++) {
                if (done[i])
                    continue;
                seen++;
                const Prim &p = prims[i];
                if (p.type == Prim::MATK)
                    break; // full barrier
                const bool light = cfg.free_perms && is_perm[i];
                // balance: a pass is HBM-bound up to ~max_heavy gates; beyond that the arithmetic is
                // the limit, so later passes (which stream the state anyway) should take the rest
                bool fits = !blk.blocked(p) && (light || heavy < cfg.max_heavy);
                if (fits && p.type == Prim::C1Q) {
                    B2_ABORT_IF(p.target >= cfg.n_local,
                                "internal: non-diagonal target on a global (rank) qubit");
                    if (!(tmask & bit(p.target))) {
                        if (grow > 0) {
                            tmask |= bit(p.target);
                            grow--;
                        } else {
                            fits = false;
                        }
                    }
                }
                if (fits) {
                    if (out)
                        out->push_back(i);
                    heavy += light ? 0 : 1;
                    n_light += light ? 1 : 0;
                    n_ops++;
                    misses = 0;
                } else {
                    blk.skip(p);
                    misses++;
                }
            }
            if (tmask_out)
                *tmask_out = tmask;
            return heavy * 64 + n_light;
        };
        uint64_t tile_mask = low_mask;
        int free_bits = B - low;
        if (prims[first].type == Prim::C1Q && !(tile_mask & bit(prims[first].target)) && free_bits > 0) {
            tile_mask |= bit(prims[first].target); // progress: the oldest pending op always runs
            free_bits--;
        }
        // Grow the tile one bit at a time by marginal gain over a look-ahead window: the candidate
        // whose addition lets the most arithmetic join the pass wins, the earliest one on ties.
        // (For layered circuits this finds the blocks of neighbouring wires whose gates of several
        // layers can run in one sweep.)
        constexpr int kWindow = 384, kMaxCand = 20;
        while (free_bits > 0 && cfg.lookahead) {
            std::vector<int> cand;
            {
                uint64_t seen_bits = tile_mask;
                int seen = 0;
                for (int i = first; i < N && seen < kWindow && static_cast<int>(cand.size()) < kMaxCand; i++) {
                    if (done[i])
                        continue;
                    seen++;
                    const Prim &p = prims[i];
                    if (p.type == Prim::MATK)
                        break;
                    if (p.type == Prim::C1Q && !(seen_bits & bit(p.target)) && p.target < cfg.n_local) {
                        seen_bits |= bit(p.target);
                        cand.push_back(p.target);
                    }
                }
            }
            if (cand.empty())
                break;
            int best = cand[0], best_score = -1;
            if (cfg.lookahead)
            for (int c : cand) {
                const int sc = select(tile_mask | bit(c), 0, kWindow, nullptr, nullptr);
                if (sc > best_score) {
                    best_score = sc;
                    best = c;
                }
            }
            tile_mask |= bit(best);
            free_bits--;
        }
        std::vector<int> chosen;
        chosen.reserve(kMaxOpsPerPass);
        // bounded look-ahead: ops further than kScanWindow pending ops away wait for a later pass
        // (skipping is always legal), which keeps scheduling linear in the circuit length
        constexpr int kScanWindow = 8192;
        select(tile_mask, free_bits, kScanWindow, &chosen, &tile_mask);
        free_bits = B - __builtin_popcountll(tile_mask);
        for (int i : chosen)
            done[i] = 1;
        B2_ASSERT(!chosen.empty());
        for (int b = 0; free_bits > 0; b++) { // pad the tile with the lowest unused local bits
            B2_ASSERT(b < std::max(cfg.n_local, cfg.n_alloc));
            if (!(tile_mask & bit(b))) {
                tile_mask |= bit(b);
                free_bits--;
            }
        }
        // The pass is built for the swizzled tile layout first; when no round keeps one of the lowest SW
        // tile bits in registers (so the lanes of a shared-memory wavefront can always enumerate the
        // bank groups through those bits) it is rebuilt for the PLAIN layout, which the load warps can
        // fill with bulk async copies of whole HBM runs (cp.async.bulk) instead of one 16-byte cp.async
        // per amplitude.
        auto make_pass = [&](bool plain) -> Pass {
        auto PH = [&](uint32_t v) { return plain ? v : phys_slot(v, B, cfg.SW, cfg.SH); };
        Pass ps;
        ps.hdr.low_bits = low;
        {
            int j = 0;
            for (int b = 0; b < 64; b++)
                if (tile_mask & bit(b))
                    ps.hdr.tile_bits[j++] = static_cast<uint8_t>(b);
            B2_ASSERT(j == B);
        }
        // a load thread's e-th copy is the 16-byte unit e * 128 + (thread): tile-local index (e * 128) << SH
        for (int e = 0; e < 64 && (e << (7 + cfg.SH)) < (1 << B); e++) {
            uint64_t off = 0;
            for (int j = 7 + cfg.SH; j < B; j++)
                if (((e << (7 + cfg.SH)) >> j) & 1)
                    off |= bit(ps.hdr.tile_bits[j]);
            ps.hdr.load_off[e] = off * (cfg.f32 ? 8u : 16u); // bytes: the load warps add it to a byte pointer
        }
        fill_tile_id_segments(ps.hdr, tile_mask, std::max(cfg.n_alloc, cfg.n_local));
        // ---- rounds, with permutation primitives folded into the address map at round boundaries
        uint32_t Mcol[kMaxTileBits]; // storage index (before the swizzle) of logical basis vector e_j
        uint32_t McolLast[kMaxTileBits] = {0}; // Mcol as the most recent round saw it
        for (int j = 0; j < B; j++)
            Mcol[j] = 1u << j;
        int n_cx = 0;
        auto tile_pos = [&](int b) {
            int j = 0;
            while (ps.hdr.tile_bits[j] != b)
                j++;
            return j;
        };
        // 0 = not absorbable, 1 = CNOT inside the tile, 2 = X toggled by CTA-uniform bits
        auto perm_class = [&](int i) {
            if (!cfg.free_perms || !is_perm[i])
                return 0;
            const Prim &p = prims[i];
            const uint64_t in_tile = p.cmask & tile_mask;
            if (in_tile == 0)
                return n_cx < kMaxCx ? 2 : 0;
            if (in_tile == p.cmask && popc(p.cmask) == 1 && p.cval == p.cmask)
                return 1;
            return 0;
        };
        // logical-space image of the permutations absorbed since the last round was built:
        // index j (as the last round's threads know it) ends up at Lmap(j) ^ (toggles that fire)
        uint32_t Lcol[kMaxTileBits];
        uint32_t cx_logical[kMaxCx];
        auto reset_logical = [&]() {
            for (int j = 0; j < B; j++)
                Lcol[j] = 1u << j;
            for (int c = 0; c < kMaxCx; c++)
                cx_logical[c] = 0;
        };
        reset_logical();
        auto absorb = [&](const Prim &p, int cls, int round_idx) {
            const int jt = tile_pos(p.target);
            if (cls == 1) {
                const int jc = tile_pos(ctz(p.cmask));
                Mcol[jc] ^= Mcol[jt];
                for (int j = 0; j < B; j++)
                    if ((Lcol[j] >> jc) & 1u)
                        Lcol[j] ^= 1u << jt;
                for (int c = 0; c < n_cx; c++)
                    if ((cx_logical[c] >> jc) & 1u)
                        cx_logical[c] ^= 1u << jt;
            } else {
                cx_logical[n_cx] = 1u << jt;
                DevCx &c = ps.hdr.cx[n_cx++];
                c.gcm = p.cmask;
                c.gcv = p.cval;
                c.vec = static_cast<uint16_t>(PH(Mcol[jt]));
                c.round = static_cast<uint16_t>(round_idx);
            }
            ps.n_absorbed++;
        };
        std::vector<int> remaining = chosen;
        int n_rounds = 0;
        while (true) {
            { // boundary: absorb every permutation that commutes back to here
                Blocker lb;
                std::vector<int> rest;
                rest.reserve(remaining.size());
                for (int i : remaining) {
                    const Prim &p = prims[i];
                    const int cls = lb.blocked(p) ? 0 : perm_class(i);
                    if (cls) {
                        absorb(p, cls, n_rounds);
                    } else {
                        lb.skip(p);
                        rest.push_back(i);
                    }
                }
                remaining.swap(rest);
            }
            if (remaining.empty())
                break;
            B2_ABORT_IF(n_rounds >= kMaxRounds, "internal: too many rounds in a pass");
            uint32_t reg_mask = 0; // over tile-local positions
            int reg_free = R;
            // Two sweeps over the remaining ops. The first only takes gates on the lane bits
            // (tile positions 0..4): doing them early keeps those bits out of the LAST round's
            // register set, which is what allows the fused store (coalesced HBM writes need the
            // lanes on the low index bits). The second sweep takes whatever else fits. An op taken
            // in the first sweep was not blocked by anything before it, so it commutes with every
            // earlier op that is still waiting.
            std::vector<int> now, later;
            now.reserve(remaining.size());
            later.reserve(remaining.size());
            std::vector<char> taken(remaining.size(), 0);
            for (int sweep = cfg.fuse_store ? 0 : 1; sweep < 2; sweep++) {
                Blocker rb;
                for (size_t r = 0; r < remaining.size(); r++) {
                    if (taken[r])
                        continue;
                    const Prim &p = prims[remaining[r]];
                    // an absorbable permutation waits for the next round boundary, where it is free
                    bool fits = !rb.blocked(p) && perm_class(remaining[r]) == 0;
                    if (fits && sweep == 0)
                        fits = p.type == Prim::C1Q && tile_pos(p.target) < low;
                    if (fits && p.type == Prim::C1Q) {
                        const int j = tile_pos(p.target);
                        if (!(reg_mask & (1u << j))) {
                            if (reg_free > 0) {
                                reg_mask |= 1u << j;
                                reg_free--;
                            } else {
                                fits = false;
                            }
                        }
                    }
                    if (fits) {
                        now.push_back(remaining[r]);
                        taken[r] = 1;
                    } else {
                        rb.skip(p);
                    }
                }
            }
            for (size_t r = 0; r < remaining.size(); r++)
                if (!taken[r])
                    later.push_back(remaining[r]);
            B2_ASSERT(!now.empty());
            // register slots in order of first use by the round's ops, then the padding bits
            std::vector<int> slot_bits;
            slot_bits.reserve(kMaxRegBits + 1);
            for (int i : now) {
                const Prim &p = prims[i];
                if (p.type != Prim::C1Q)
                    continue;
                const int j = tile_pos(p.target);
                if (std::find(slot_bits.begin(), slot_bits.end(), j) == slot_bits.end())
                    slot_bits.push_back(j);
            }
            for (int j = B - 1; reg_free > 0; j--) { // pad with the highest unused tile bits
                B2_ASSERT(j >= 0);
                if (!(reg_mask & (1u << j))) {
                    reg_mask |= 1u << j;
                    reg_free--;
                    slot_bits.push_back(j);
                }
            }
            B2_ASSERT(static_cast<int>(slot_bits.size()) == R);
            // dense round: every op is an uncontrolled, untagged 2x2 on its own register slot
            bool dense = static_cast<int>(now.size()) <= R;
            for (size_t k = 0; dense && k < now.size(); k++) {
                const Prim &p = prims[now[k]];
                dense = p.type == Prim::C1Q && p.cmask == 0 && p.tag < 0 &&
                        tile_pos(p.target) == slot_bits[k];
            }
            ps.hdr.round_kind[n_rounds] = dense ? static_cast<uint8_t>(now.size()) : 0;
            if (dense && factor && now.size() >= 2 && n_rounds < kMaxDense) {
                std::vector<Factored> fs(now.size());
                bool ok = true;
                for (size_t k = 0; ok && k < now.size(); k++)
                    ok = factor_2x2(prims[now[k]].m, fs[k]);
                if (ok) {
                    DevDense dd;
                    if (cfg.f32)
                        fill_dense<float>(dd, fs);
                    else
                        fill_dense<double>(dd, fs);
                    ps.hdr.round_kind[n_rounds] = static_cast<uint8_t>(8 + now.size());
                    if (ps.dense.size() <= static_cast<size_t>(n_rounds))
                        ps.dense.resize(n_rounds + 1, DevDense{});
                    ps.dense[n_rounds] = dd;
                }
            }
            uint8_t *rbits = ps.hdr.round_regbits[n_rounds];
            {
                std::vector<uint32_t> free_cols; // ((1 << j) << 16) | storage column of free bit j
                for (int s = 0; s < R; s++) {
                    rbits[s] = static_cast<uint8_t>(slot_bits[s]);
                    ps.hdr.round_poff[n_rounds][s] =
                        static_cast<uint16_t>(PH(Mcol[slot_bits[s]]));
                }
                for (int j = 0; j < B; j++)
                    if (!(reg_mask & (1u << j)))
                        free_cols.push_back(((1u << j) << 16) | PH(Mcol[j]));
                B2_ASSERT(free_cols.size() <= static_cast<size_t>(kMaxFreeBits));
                // Bank-conflict-free gathers: the lanes that share one shared-memory wavefront
                // (8 x 16 B or 16 x 8 B) differ in the lowest SW thread-id bits, so give those
                // bits free tile bits whose storage columns are linearly independent (over GF(2))
                // in their low SW bits = they enumerate all 2^SW bank groups.
                const uint32_t bank_mask = (1u << cfg.SW) - 1u;
                std::vector<uint32_t> order, rest, basis;
                for (uint32_t fc : free_cols) {
                    uint32_t v = fc & bank_mask;
                    for (uint32_t b : basis)
                        v = std::min(v, v ^ b);
                    if (v != 0 && static_cast<int>(order.size()) < cfg.SW) {
                        basis.push_back(v);
                        std::sort(basis.rbegin(), basis.rend());
                        order.push_back(fc);
                    } else {
                        rest.push_back(fc);
                    }
                }
                order.insert(order.end(), rest.begin(), rest.end());
                for (size_t k = 0; k < order.size(); k++)
                    ps.hdr.round_col[n_rounds][k] = order[k];
            }
            for (int j = 0; j < B; j++)
                McolLast[j] = Mcol[j];
            ps.hdr.round_begin[n_rounds] = static_cast<uint16_t>(ps.ops.size());
            for (int i : now) {
                ps.ops.push_back(make_devop(prims[i], ps.hdr.tile_bits, B, rbits, R, -1));
                ps.tags.push_back(prims[i].tag);
            }
            n_rounds++;
            remaining.swap(later);
            reset_logical();
        }
        if (n_rounds >= 1 && cfg.fuse_store) {
            const int last = n_rounds - 1;
            uint32_t reg_mask = 0;
            for (int k = 0; k < R; k++)
                reg_mask |= 1u << ps.hdr.round_regbits[last][k];
            auto spread = [&](uint32_t v) {
                uint64_t g = 0;
                for (int j = 0; j < B; j++)
                    if ((v >> j) & 1u)
                        g |= bit(ps.hdr.tile_bits[j]);
                return g;
            };
            // thread-id bit k of the last round <-> free tile bit lanes[k]; fills the store maps
            auto commit = [&](const std::vector<int> &lanes, int mode) {
                for (size_t k = 0; k < lanes.size(); k++) {
                    const int j = lanes[k];
                    ps.hdr.round_col[last][k] =
                        ((1u << j) << 16) | PH(McolLast[j]);
                    ps.hdr.store_free[k] = spread(Lcol[j]);
                }
                for (int k = 0; k < R; k++) {
                    const uint32_t v = Lcol[ps.hdr.round_regbits[last][k]];
                    ps.hdr.store_reg[k] = spread(v);
                }
                for (int c = 0; c < n_cx; c++) {
                    const uint32_t v = ps.hdr.cx[c].round == n_rounds ? cx_logical[c] : 0;
                    ps.hdr.store_cx[c] = spread(v);
                }
                ps.hdr.fused_store = static_cast<uint8_t>(mode);
            };
            std::vector<int> free_js;
            for (int j = 0; j < B; j++)
                if (!(reg_mask & (1u << j)))
                    free_js.push_back(j);
            {
                // Direct store from registers: possible when `low` thread-id bits of the last round can
                // be given free tile bits whose images under the trailing permutations stay inside the
                // contiguous low index bits and span them -- then the 32 lanes of a warp (the five
                // lowest thread-id bits: those `low` plus any 5 - low other free bits) always cover
                // whole runs of 2^low amplitudes.
                std::vector<int> lanes, others;
                std::vector<uint32_t> basis;
                const int n_lane = std::min(low, 5);
                for (int j : free_js) {
                    uint32_t v = Lcol[j];
                    bool ok = v < (1u << n_lane) && static_cast<int>(lanes.size()) < n_lane;
                    if (ok) {
                        for (uint32_t b : basis)
                            v = std::min(v, v ^ b);
                        ok = v != 0;
                    }
                    if (ok) {
                        basis.push_back(v);
                        std::sort(basis.rbegin(), basis.rend());
                        lanes.push_back(j);
                    } else {
                        others.push_back(j);
                    }
                }
                if (static_cast<int>(lanes.size()) == n_lane) {
                    lanes.insert(lanes.end(), others.begin(), others.end());
                    commit(lanes, 1);
                }
            }
        }
        for (int j = 0; j < B; j++)
            ps.hdr.final_col[j] = static_cast<uint16_t>(PH(Mcol[j]));
        ps.hdr.n_cx = n_cx;
        ps.hdr.round_begin[n_rounds] = static_cast<uint16_t>(ps.ops.size());
        ps.hdr.n_rounds = n_rounds;
        ps.hdr.n_ops = static_cast<int32_t>(ps.ops.size());
        ps.hdr.plain_layout = plain ? 1 : 0;
        {
            int run_bits = 0;
            while (run_bits < B && ps.hdr.tile_bits[run_bits] == run_bits)
                run_bits++;
            ps.hdr.bulk_run_bits = static_cast<uint8_t>(std::max(run_bits, low));
        }
        return ps;
        };
        Pass ps = make_pass(false);
        if (cfg.bulk && low >= cfg.SW) {
            // A bulk copy costs the SM's copy engine ~46 cycles whatever its size (B300_MICROARCH.md,
            // "TMA service/SM"), so 512-byte runs cap a tile at ~11 B/cycle -- measured 10 % slower than
            // the per-amplitude cp.async loader (profiles/r2_tile_bulk_ab.md). Bulk loads are used where
            // the tile's low bits form runs of at least cfg.bulk_min_run_bits amplitudes (>= 4 KiB).
            int run_bits = 0;
            while (run_bits < B && ps.hdr.tile_bits[run_bits] == run_bits)
                run_bits++;
            bool eligible = ps.hdr.n_rounds >= 1 && run_bits >= cfg.bulk_min_run_bits;
            for (int rd = 0; rd < ps.hdr.n_rounds && eligible; rd++)
                for (int k = 0; k < R; k++)
                    eligible = eligible && ps.hdr.round_regbits[rd][k] >= cfg.SW;
            if (eligible)
                ps = make_pass(true);
        }
        passes.push_back(std::move(ps));
    }
    return passes;
}

} // namespace

} // namespace b2sv
