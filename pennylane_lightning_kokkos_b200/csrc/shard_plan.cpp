// b2sv: planner for sharded states (see shard_plan.hpp).
#include "shard_plan.hpp"

#include <algorithm>

namespace b2sv {

namespace {
uint64_t map_mask(uint64_t m, const std::vector<int> &l2p) {
    uint64_t r = 0;
    while (m) {
        const int q = __builtin_ctzll(m);
        r |= bit(l2p[q]);
        m &= m - 1;
    }
    return r;
}
} // namespace

Prim prim_to_physical(const Prim &p, const std::vector<int> &l2p) {
    Prim q = p;
    if (p.type == Prim::C1Q)
        q.target = l2p[p.target];
    q.cmask = map_mask(p.cmask, l2p);
    q.cval = map_mask(p.cval, l2p);
    q.pmask = map_mask(p.pmask, l2p);
    for (int &b : q.bits)
        b = l2p[b];
    return q;
}

std::vector<ShardStep> plan_sharded(std::vector<Prim> pending, std::vector<int> &l2p,
                                    const ShardPlanConfig &cfg) {
    const int n = cfg.n, n_local = cfg.n_local, g = n - n_local;
    std::vector<ShardStep> steps;
    while (!pending.empty()) {
        // ---- everything that can run in the current layout, in program order
        ShardStep run;
        std::vector<Prim> rest;
        uint64_t T = 0, D = 0; // logical bits touched non-diagonally / diagonally by skipped ops
        for (const Prim &p : pending) {
            const uint64_t tm = p.target_mask(), dm = p.support() & ~tm;
            const bool blocked = (tm & (T | D)) || (dm & T);
            const bool local = (map_mask(tm, l2p) >> n_local) == 0;
            if (!blocked && local) {
                run.prims.push_back(prim_to_physical(p, l2p));
            } else {
                T |= tm;
                D |= dm;
                rest.push_back(p);
            }
        }
        if (!run.prims.empty())
            steps.push_back(std::move(run));
        if (rest.empty())
            break;
        // ---- first / next non-diagonal use of every logical qubit among the waiting ops
        const size_t never = rest.size() + 1;
        std::vector<size_t> next_use(n, never);
        for (size_t i = rest.size(); i-- > 0;) {
            uint64_t tm = rest[i].target_mask();
            while (tm) {
                next_use[__builtin_ctzll(tm)] = i;
                tm &= tm - 1;
            }
        }
        // global qubits that will be needed, soonest first
        std::vector<int> need;
        for (int q = 0; q < n; q++)
            if (l2p[q] >= n_local && next_use[q] != never)
                need.push_back(q);
        std::sort(need.begin(), need.end(), [&](int a, int b) { return next_use[a] < next_use[b]; });
        B2_ASSERT(!need.empty());
        if (!cfg.batch)
            need.resize(1);
        B2_ASSERT(static_cast<int>(need.size()) <= g);
        uint64_t need_mask = 0;
        for (int q : need)
            need_mask |= bit(q);
        // victims: local qubits, farthest next use first; the low positions only as a last resort
        std::vector<int> cand;
        for (int min_pos : {cfg.min_victim_pos, 0}) {
            cand.clear();
            for (int o = 0; o < n; o++)
                if (l2p[o] < n_local && l2p[o] >= min_pos && l2p[o] < cfg.max_victim_pos &&
                    !((need_mask >> o) & 1))
                    cand.push_back(o);
            if (cand.size() >= need.size())
                break;
        }
        B2_ABORT_IF(cand.size() < 1, "operation acts on more qubits than one shard holds");
        std::sort(cand.begin(), cand.end(), [&](int a, int b) {
            if (next_use[a] != next_use[b])
                return next_use[a] > next_use[b];
            return l2p[a] > l2p[b];
        });
        ShardStep ex;
        ex.is_exchange = true;
        for (size_t i = 0; i < need.size() && i < cand.size(); i++) {
            const int q = need[i], v = cand[i];
            // beyond the first (which the frontier is waiting for): only worth it when the incoming
            // qubit is needed before the outgoing one
            if (i > 0 && next_use[v] <= next_use[q])
                break;
            ex.swaps.emplace_back(l2p[q], l2p[v]);
            std::swap(l2p[q], l2p[v]);
        }
        steps.push_back(std::move(ex));
        pending.swap(rest);
    }
    return steps;
}

std::vector<ShardStep> plan_normalize(std::vector<int> &l2p, const ShardPlanConfig &cfg) {
    const int n = cfg.n, n_local = cfg.n_local;
    std::vector<ShardStep> steps;
    std::vector<int> p2l(n);
    auto refresh = [&]() {
        for (int q = 0; q < n; q++)
            p2l[l2p[q]] = q;
    };
    refresh();
    // 1. every logical rank bit that sits on a WRONG rank position comes into the shard first
    //    (an exchange pairs a rank position with a local one; rank<->rank moves take two steps)
    {
        ShardStep ex;
        ex.is_exchange = true;
        std::vector<char> used(n_local, 0);
        for (int P = n_local; P < n; P++) {
            if (l2p[P] == P || l2p[P] < n_local)
                continue;
            int victim = -1; // a local position holding a logical LOCAL qubit, high positions first
            for (int pos = n_local - 1; pos >= 0 && victim < 0; pos--)
                if (!used[pos] && p2l[pos] < n_local)
                    victim = pos;
            B2_ASSERT(victim >= 0);
            used[victim] = 1;
            ex.swaps.emplace_back(l2p[P], victim);
        }
        if (!ex.swaps.empty()) {
            for (const auto &s : ex.swaps) {
                const int qa = p2l[s.first], qb = p2l[s.second];
                std::swap(l2p[qa], l2p[qb]);
            }
            refresh();
            steps.push_back(std::move(ex));
        }
    }
    // 2. now every misplaced logical rank bit is local: one exchange puts them all home
    {
        ShardStep ex;
        ex.is_exchange = true;
        for (int P = n_local; P < n; P++)
            if (l2p[P] != P) {
                B2_ASSERT(l2p[P] < n_local);
                ex.swaps.emplace_back(P, l2p[P]);
            }
        if (!ex.swaps.empty()) {
            for (const auto &s : ex.swaps) {
                const int qa = p2l[s.first], qb = p2l[s.second];
                std::swap(l2p[qa], l2p[qb]);
            }
            refresh();
            steps.push_back(std::move(ex));
        }
    }
    return steps;
}

} // namespace b2sv
