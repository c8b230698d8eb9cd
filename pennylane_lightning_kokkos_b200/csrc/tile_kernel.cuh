// b2sv: the tile executor -- ONE kernel applies every gate of a fused pass in a single HBM sweep.
//
// Replaces the reference's one-gate-per-sweep functor dispatch
// (reference StateVectorKokkos.hpp:807-824 applyGateFunctor -> GateFunctors.hpp, one
// Kokkos::parallel_for over 2^(n-k) per gate).
//
// Execution model (one persistent CTA per SM: two worker groups + one load warpgroup, three rotating
// tile buffers):
//   * a tile = 2^B amplitudes (B = 12 for complex128 -> 64 KiB) whose indices differ in the pass's
//     B tile bits; the CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
//   * the load warps stream tile k into the next free buffer (cp.async = LDGSTS with a
//     per-amplitude XOR-swizzled destination, completion on an mbarrier) and publish the tile's
//     facts (base index, conditional address toggles) next to it;
//   * each worker group (GT threads) owns every second tile of the CTA: it waits for the tile to
//     land, runs all rounds of the pass on it in registers and writes the last round straight to
//     HBM in place (or, when the address map does not allow coalesced stores, through a store phase
//     from shared memory); the buffer goes back to the load warps right after the last gather.
//     So while two tiles are being computed a third is always in flight;
//   * a round: every thread gathers the 2^R amplitudes whose indices differ only in the round's R
//     "register bits", applies all ops of the round in registers, scatters back in place. Rounds of
//     uncontrolled 2x2s run as straight-line code, in factored form (dense_factored) when the
//     scheduler could factor every gate; anything else goes through the per-op interpreter;
//   * the kernel is compiled in four variants per dtype that differ in which round bodies exist,
//     the launcher (launch_tile_pass_t) picks the leanest one that covers the pass.
// HBM traffic is exactly one read + one write of the state per pass, whatever the number of gates.
//
// Shared-memory layout: amplitude i of the tile lives at slot phys(i) = i ^ (fold(i) << SH) where fold
// XORs the higher 3-bit groups of i into bits SH .. SH+2 (SH = 0 for complex128, 1 for complex64: the
// two complex64 of a 16-byte unit stay together, so both types are loaded with 16-byte cp.async.cg).
// phys is GF(2)-linear, so phys(base | off) = phys(base) ^ phys(off), and any 8 consecutive lanes hit
// distinct 16 B bank groups for every choice of register bits.
#pragma once
#include "schedule.hpp"

#include <algorithm>
#include <cstdlib>
#include <cuda_runtime.h>

namespace b2sv {

template <typename real> struct AmpT;
template <> struct AmpT<double> {
    using type = double2;
};
template <> struct AmpT<float> {
    using type = float2;
};

template <int B, int SW, int SH = 0> __host__ __device__ __forceinline__ constexpr uint32_t phys(uint32_t i) {
    constexpr int FW = SW - SH;
    uint32_t f = 0;
    for (int s = SH + FW; s < B; s += FW)
        f ^= (i >> s);
    return i ^ ((f & ((1u << FW) - 1u)) << SH);
}

// ---- async-copy / mbarrier / named-barrier primitives ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16 bytes global -> shared (L1-bypassing), destination given as a shared-window address
template <int BYTES> __device__ __forceinline__ void cp_async_to(uint32_t dst, const void *src) {
    static_assert(BYTES == 16, "the tile loader copies 16-byte units");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival from this thread once all its cp.async issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *mbar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(mbar))
                 : "memory");
}
// Bulk async copy (TMA engine, non-tensor form): `bytes` contiguous bytes global -> shared, completion
// counted in bytes on the mbarrier. Addresses and size are multiples of 16.
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
// one arrival that also announces `bytes` of bulk-copy traffic to wait for
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *mbar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(
                     smem_u32(mbar)),
                 "r"(bytes)
                 : "memory");
}
// generic-proxy accesses of this thread (tile gathers / scatters) before async-proxy writes that follow
// the hand-off (the next bulk copy into the same buffer)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(
                     smem_u32(mbar))
                 : "memory");
}
// Polling must not steal issue slots from the arithmetic warps of the same SM sub-partition: the
// try_wait carries a suspend-time hint and a missed poll backs off with nanosleep.
// A lost hand-off must surface as a CUDA error, never as a hung GPU: the wait gives up after
// kMbarTimeoutNs of wall time (%globaltimer, looked at every 4096 missed polls), not after a number of
// polls whose duration depends on the sleep granularity.
constexpr unsigned long long kMbarTimeoutNs = 10ull * 1000 * 1000 * 1000;
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    const uint32_t a = smem_u32(mbar);
    uint32_t ok;
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (true) {
        asm volatile("{\n .reg .pred p;\n"
                     " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                     " selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(a), "r"(parity), "r"(static_cast<uint32_t>(SLEEP_NS * 4))
                     : "memory");
        if (ok)
            break;
        __nanosleep(SLEEP_NS);
        if ((++spins & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0)
                t0 = now;
            else if (now - t0 > kMbarTimeoutNs)
                __trap();
        }
    }
}
// one lane polls, the rest of the warp parks at the warp barrier (no 32-wide spinning)
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_warp(uint64_t *mbar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0)
        mbar_wait<SLEEP_NS>(mbar, parity);
    __syncwarp();
}
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- register-level op bodies -------------------------------------------------------------------
template <int TS, int NS, typename amp_t, typename real>
__device__ __forceinline__ void op_general(amp_t (&a)[NS], const real (&m)[8], uint32_t act,
                                           bool pred) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s & (1 << TS))
            continue;
        constexpr int dummy = 0;
        (void)dummy;
        const int s1 = s | (1 << TS);
        if (pred && ((act >> s) & 1u)) {
            const amp_t v0 = a[s], v1 = a[s1];
            amp_t r0, r1;
            r0.x = m[0] * v0.x - m[1] * v0.y + m[2] * v1.x - m[3] * v1.y;
            r0.y = m[0] * v0.y + m[1] * v0.x + m[2] * v1.y + m[3] * v1.x;
            r1.x = m[4] * v0.x - m[5] * v0.y + m[6] * v1.x - m[7] * v1.y;
            r1.y = m[4] * v0.y + m[5] * v0.x + m[6] * v1.y + m[7] * v1.x;
            a[s] = r0;
            a[s1] = r1;
        }
    }
}
template <int TS, int NS, typename amp_t, typename real>
__device__ __forceinline__ void op_real(amp_t (&a)[NS], const real (&m)[8], uint32_t act,
                                        bool pred) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s & (1 << TS))
            continue;
        const int s1 = s | (1 << TS);
        if (pred && ((act >> s) & 1u)) {
            const amp_t v0 = a[s], v1 = a[s1];
            amp_t r0, r1;
            r0.x = m[0] * v0.x + m[2] * v1.x;
            r0.y = m[0] * v0.y + m[2] * v1.y;
            r1.x = m[4] * v0.x + m[6] * v1.x;
            r1.y = m[4] * v0.y + m[6] * v1.y;
            a[s] = r0;
            a[s1] = r1;
        }
    }
}
template <int TS, int NS, typename amp_t>
__device__ __forceinline__ void op_perm(amp_t (&a)[NS], uint32_t act, bool pred) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s & (1 << TS))
            continue;
        const int s1 = s | (1 << TS);
        if (pred && ((act >> s) & 1u)) {
            const amp_t t = a[s];
            a[s] = a[s1];
            a[s1] = t;
        }
    }
}
template <int NS, typename amp_t, typename real>
__device__ __forceinline__ void op_diag(amp_t (&a)[NS], const real (&m)[8], uint32_t act,
                                        uint32_t par, bool odd_base, bool pred) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (pred && ((act >> s) & 1u)) {
            const bool odd = odd_base ^ (((par >> s) & 1u) != 0);
            const real pr = odd ? m[2] : m[0];
            const real pi = odd ? m[3] : m[1];
            const amp_t v = a[s];
            amp_t r;
            r.x = pr * v.x - pi * v.y;
            r.y = pr * v.y + pi * v.x;
            a[s] = r;
        }
    }
}

// ---- fast paths: no control of any kind, so no predicate and full instruction-level parallelism
template <int TS, int NS, typename amp_t, typename real>
__device__ __forceinline__ void fast_general(amp_t (&a)[NS], const real (&m)[8]) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s & (1 << TS))
            continue;
        const int s1 = s | (1 << TS);
        const amp_t v0 = a[s], v1 = a[s1];
        amp_t r0, r1;
        r0.x = m[0] * v0.x - m[1] * v0.y + m[2] * v1.x - m[3] * v1.y;
        r0.y = m[0] * v0.y + m[1] * v0.x + m[2] * v1.y + m[3] * v1.x;
        r1.x = m[4] * v0.x - m[5] * v0.y + m[6] * v1.x - m[7] * v1.y;
        r1.y = m[4] * v0.y + m[5] * v0.x + m[6] * v1.y + m[7] * v1.x;
        a[s] = r0;
        a[s1] = r1;
    }
}
template <int TS, int NS, typename amp_t, typename real>
__device__ __forceinline__ void fast_real(amp_t (&a)[NS], const real (&m)[8]) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s & (1 << TS))
            continue;
        const int s1 = s | (1 << TS);
        const amp_t v0 = a[s], v1 = a[s1];
        amp_t r0, r1;
        r0.x = m[0] * v0.x + m[2] * v1.x;
        r0.y = m[0] * v0.y + m[2] * v1.y;
        r1.x = m[4] * v0.x + m[6] * v1.x;
        r1.y = m[4] * v0.y + m[6] * v1.y;
        a[s] = r0;
        a[s1] = r1;
    }
}

template <int R, int NS, typename amp_t, typename real>
__device__ __forceinline__ void run_op(amp_t (&a)[NS], const DevOp &op, uint64_t tile_base,
                                       uint32_t base_local) {
    real m[8];
    {
        const double2 *mp = reinterpret_cast<const double2 *>(op.m);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 t = mp[i];
            m[2 * i] = static_cast<real>(t.x);
            m[2 * i + 1] = static_cast<real>(t.y);
        }
    }
    const int kind = op.kind & OPF_KIND_MASK;
    const int ts = op.tslot;
    if ((op.kind & OPF_UNCOND) && kind <= KIND_REAL) {
        // code = kind * 8 + ts; one flat switch keeps the hot bodies contiguous
        switch (kind * 8 + ts) {
#define B2_FAST(TS)                                                                             \
    case TS:                                                                                    \
        if constexpr (TS < R)                                                                   \
            fast_general<TS, NS>(a, m);                                                         \
        break;                                                                                  \
    case 8 + TS:                                                                                \
        if constexpr (TS < R)                                                                   \
            fast_real<TS, NS>(a, m);                                                            \
        break;
            B2_FAST(0)
            B2_FAST(1)
            B2_FAST(2)
            B2_FAST(3)
            B2_FAST(4)
#undef B2_FAST
        default:
            break;
        }
        return;
    }
    if ((tile_base & op.gcm) != op.gcv)
        return; // CTA-uniform: the whole tile fails the control
    const bool pred = (base_local & op.lcm) == op.lcv;
    const uint32_t act = op.slot_act;
    if (kind == KIND_DIAG) {
        const bool odd =
            ((__popcll(tile_base & op.gpm) + __popc(base_local & op.lpm)) & 1) != 0;
        op_diag<NS>(a, m, act, op.slot_par, odd, pred);
        return;
    }
    switch (ts) {
#define B2_CASE(TS)                                                                             \
    case TS:                                                                                    \
        if constexpr (TS < R) {                                                                 \
            if (kind == KIND_GENERAL)                                                           \
                op_general<TS, NS>(a, m, act, pred);                                            \
            else if (kind == KIND_REAL)                                                         \
                op_real<TS, NS>(a, m, act, pred);                                               \
            else                                                                                \
                op_perm<TS, NS>(a, act, pred);                                                  \
        }                                                                                       \
        break;
        B2_CASE(0)
        B2_CASE(1)
        B2_CASE(2)
        B2_CASE(3)
        B2_CASE(4)
#undef B2_CASE
    default:
        break;
    }
}

// G uncontrolled 2x2 gates, gate k on register slot k (k < R), fully unrolled.
template <int G, int R, int NS, typename amp_t, typename real>
__device__ __forceinline__ void dense_round(amp_t (&a)[NS], const DevOp *ops) {
#pragma unroll
    for (int k = 0; k < G; k++) {
        if constexpr (true) {
            if (k >= R)
                break;
        }
        real m[8];
        const double2 *mp = reinterpret_cast<const double2 *>(ops[k].m);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 t = mp[i];
            m[2 * i] = static_cast<real>(t.x);
            m[2 * i + 1] = static_cast<real>(t.y);
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (s & (1 << k))
                continue;
            const int s1 = s | (1 << k);
            const amp_t v0 = a[s], v1 = a[s1];
            amp_t r0, r1;
            r0.x = m[0] * v0.x - m[1] * v0.y + m[2] * v1.x - m[3] * v1.y;
            r0.y = m[0] * v0.y + m[1] * v0.x + m[2] * v1.y + m[3] * v1.x;
            r1.x = m[4] * v0.x - m[5] * v0.y + m[6] * v1.x - m[7] * v1.y;
            r1.y = m[4] * v0.y + m[5] * v0.x + m[6] * v1.y + m[7] * v1.x;
            a[s] = r0;
            a[s1] = r1;
        }
    }
}
// The same round in factored form (DevDense): per-slot phase, G real shears (one FMA per real
// number and gate), per-slot scale. 4 + 2 G + 4 multiply-adds per amplitude instead of 8 G.
template <int G, int R, int NS, typename amp_t, typename real>
__device__ __forceinline__ void dense_factored(amp_t (&a)[NS], const DevDense &dd) {
    constexpr int GG = G < R ? G : R;
    constexpr int GM = (1 << GG) - 1;
    const amp_t *pre = reinterpret_cast<const amp_t *>(dd.pre);
    const amp_t *post = reinterpret_cast<const amp_t *>(dd.post);
    const real *t = reinterpret_cast<const real *>(dd.t);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if ((s & GM) == 0)
            continue; // pre[0] = 1
        const amp_t p = pre[s & GM];
        const amp_t v = a[s];
        a[s].x = p.x * v.x - p.y * v.y;
        a[s].y = p.x * v.y + p.y * v.x;
    }
#pragma unroll
    for (int k = 0; k < GG; k++) {
        const real t01 = t[2 * k], t10 = t[2 * k + 1];
        if constexpr (sizeof(real) == 4) {
            // complex64: a real shear acts on (re, im) alike -- one packed FFMA2 (sm_100 fma.rn.f32x2,
            // each half rounded like a scalar FMA) per amplitude instead of two FFMA
            const float2 T01 = make_float2(t01, t01), T10 = make_float2(t10, t10);
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (s & (1 << k))
                    continue;
                const int s1 = s | (1 << k);
                const float2 v0 = a[s], v1 = a[s1];
                a[s] = __ffma2_rn(T01, v1, v0);
                a[s1] = __ffma2_rn(T10, v0, v1);
            }
        } else {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (s & (1 << k))
                continue;
            const int s1 = s | (1 << k);
            const amp_t v0 = a[s], v1 = a[s1];
            a[s].x = fma(t01, v1.x, v0.x);
            a[s].y = fma(t01, v1.y, v0.y);
            a[s1].x = fma(t10, v0.x, v1.x);
            a[s1].y = fma(t10, v0.y, v1.y);
        }
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const amp_t p = post[s & GM];
        const amp_t v = a[s];
        a[s].x = p.x * v.x - p.y * v.y;
        a[s].y = p.x * v.y + p.y * v.x;
    }
}
// pb / poff: byte offsets from the start of the tile buffers (buffer offset included in pb)
template <int R, int NS, typename amp_t>
__device__ __forceinline__ void scatter_round(unsigned char *tiles_raw, const amp_t (&a)[NS], uint32_t pb,
                                              const uint32_t (&poff)[R]) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        uint32_t x = pb;
#pragma unroll
        for (int c = 0; c < R; c++)
            if (s & (1 << c))
                x ^= poff[c];
        *reinterpret_cast<amp_t *>(tiles_raw + x) = a[s];
    }
}

// last round of a pass: registers -> HBM directly. `sbase` = global index of this thread's register
// slot 0: the tile's part (base index ^ the conditional toggles that fire for it, published by the load
// warps) ^ the thread's part (tabulated once per pass); the slots are XOR-combinations of the uniform
// store_reg offsets away.
template <int R, int NS, typename amp_t>
__device__ __forceinline__ void store_round(amp_t *__restrict__ state, const amp_t (&a)[NS],
                                            const DevPassHeader &ph, uint64_t sbase) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
        uint64_t x = sbase;
#pragma unroll
        for (int c = 0; c < R; c++)
            if (s & (1 << c))
                x ^= ph.store_reg[c];
        state[x] = a[s];
    }
}
// mode: 0 = scatter back in place, 1 = registers -> HBM
template <int R, int NS, typename amp_t>
__device__ __forceinline__ void finish_round(int mode, amp_t *__restrict__ state, unsigned char *tiles_raw,
                                             const amp_t (&a)[NS], uint32_t pb,
                                             const uint32_t (&poff)[R], const DevPassHeader &ph,
                                             uint64_t sbase) {
    if (mode == 1)
        store_round<R, NS>(state, a, ph, sbase);
    else
        scatter_round<R, NS>(tiles_raw, a, pb, poff);
}

// Factored round whose constants sit at a compile-time offset of the kernel parameter (round RD):
// ptxas reads them through the uniform datapath (LDCU -> DFMA R, R, UR, R), no vector registers.
template <int RD, int R, int NS, typename amp_t, typename real>
__device__ __forceinline__ bool factored_round_fixed(int kind, amp_t (&a)[NS], const PassParams &pp,
                                                     int mode, amp_t *__restrict__ state,
                                                     unsigned char *tiles_raw, uint32_t pb,
                                                     const uint32_t (&poff)[R], uint64_t sbase) {
    switch (kind) {
#define B2_FIX(G)                                                                               \
    case 8 + G:                                                                                 \
        if constexpr (G <= R) {                                                                 \
            dense_factored<G, R, NS, amp_t, real>(a, pp.dense[RD]);                             \
            finish_round<R, NS>(mode, state, tiles_raw, a, pb, poff, pp.hdr, sbase);            \
            return true;                                                                        \
        }                                                                                       \
        break;
        B2_FIX(2)
        B2_FIX(3)
        B2_FIX(4)
        B2_FIX(5)
#undef B2_FIX
    default:
        break;
    }
    return false;
}

// ---- the kernel ---------------------------------------------------------------------------------
// The pass descriptor travels as a __grid_constant__ kernel parameter (constant bank): no upload,
// and the dense rounds read their gate matrices through uniform constant loads instead of holding
// them in 16 vector registers per thread.
// helper warps: one warpgroup that does nothing but stream tiles into shared memory (cp.async)
constexpr int kLoadThreads = 128;
constexpr int kMinLow = 3; // smallest B2SV_TILE_LOW the row-offset table is sized for

// Optional phase timers (B2SV_TILE_PROF=1): cycles summed over the lead thread of every worker group
// / load warpgroup of every CTA. [0] workers waiting for a tile, [1] workers busy on tiles,
// [2] load warps waiting for a free buffer, [3] load warps issuing copies, [4] tiles, [5] CTA lifetime,
// [6] workers in the last (fused-store) round of a tile, [7] workers in the other rounds,
// [10] workers between tiles (prologue), [11] load warps issuing one tile.
static __device__ unsigned long long g_tile_prof[16];

// per-buffer facts about the tile it holds, written by the load warps ahead of the workers
struct TileInfo {
    uint64_t tb;                    // index of the tile's first amplitude (tile bits cleared)
    uint64_t store_base;            // tb ^ the conditional toggles of the fused store that fire for this tile
    uint32_t xoff[kMaxRounds + 1];  // CTA-uniform address toggles visible from round r on
    uint32_t pad_;
};
constexpr int kAccRounds = 4; // rounds whose per-thread gather bases are tabulated in shared memory

// FACT / INTERP / DENSEK: the kernel contains the factored-round bodies / the per-op interpreter / the
// unfactored dense rounds of 2..5 gates (one-gate dense rounds are in every variant). A pass is
// launched on the leanest variant that covers its rounds: code the pass never runs would still cost
// it registers (ptxas allocates for the union of all paths of the round loop).
// BULK: the pass uses the plain tile layout and the load warps fill it with bulk async copies
// (cp.async.bulk); a template parameter so that neither path costs the other registers.
template <typename real, int B, int R, int GT, int NG, int NB, bool FACT, bool INTERP, bool DENSEK,
          bool PROF, bool BULK>
__global__ void __launch_bounds__(NG * GT + kLoadThreads, 1)
    tile_exec_kernel(typename AmpT<real>::type *__restrict__ state,
                     const __grid_constant__ PassParams pp, uint64_t rank_bits,
                     uint32_t n_tiles, uint32_t flags) {
    constexpr int kTileBuffers = NB;
    // Lockstep vs free-running worker groups. With one round per tile there is no barrier left in the
    // loop, and letting the eight warps of a group drift apart costs HBM locality (measured: 6.6 vs
    // 7.7 ms per 30-qubit sweep); with two or more rounds the free-running form wins or ties.
    // flags bit 0 forces lockstep, bit 1 forces free-running (experiments).
    // (Variants with the op interpreter always run in lockstep: carrying both release protocols through
    // the interpreter's loop costs ptxas ~800 bytes of spills per thread.)
    const bool gsync =
        INTERP ? true : ((flags & 2u) ? false : ((flags & 1u) != 0 || pp.hdr.n_rounds <= 1));
    using amp_t = typename AmpT<real>::type;
    constexpr int SW = (sizeof(amp_t) == 16) ? 3 : 4; // log2(amplitudes per 128 B)
    constexpr int SH = (sizeof(amp_t) == 16) ? 0 : 1; // log2(amplitudes per 16 B): not swizzled
    constexpr int NS = 1 << R;
    constexpr int TILE = 1 << B;
    constexpr int NF = B - R; // non-register tile bits = thread-id bits within a group
    constexpr int NTHREADS = NG * GT + kLoadThreads;
    static_assert((1 << NF) == GT, "one register group per thread");
    static_assert(NF <= kMaxFreeBits, "too many thread-id bits");
    constexpr int EPT = TILE / GT; // amplitudes per thread in the store phase (= NS)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    amp_t *tiles = reinterpret_cast<amp_t *>(smem_raw); // kTileBuffers buffers of TILE amplitudes
    DevOp *sops = reinterpret_cast<DevOp *>(smem_raw + kTileBuffers * sizeof(amp_t) * TILE);
    uint64_t *rowoff = reinterpret_cast<uint64_t *>(sops + kMaxOpsPerPass);
    // per-thread constants of the pass (the same for both worker groups): the thread part of the fused
    // store's global index, and (base << 16 | storage slot) of the thread's register group per round
    uint64_t *sfree = rowoff + (size_t(1) << (B - kMinLow));
    uint32_t *acc_tab = reinterpret_cast<uint32_t *>(sfree + GT);
    __shared__ DevPassHeader hdr;
    __shared__ TileInfo tinfo[kTileBuffers];
    __shared__ __align__(8) uint64_t full[kTileBuffers], empty[kTileBuffers];

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(&pp.hdr);
        uint4 *dst = reinterpret_cast<uint4 *>(&hdr);
        for (int i = threadIdx.x; i < static_cast<int>(sizeof(DevPassHeader) / 16); i += NTHREADS)
            dst[i] = src[i];
        const int n_ops = pp.hdr.n_ops;
        const uint4 *osrc = reinterpret_cast<const uint4 *>(pp.ops);
        uint4 *odst = reinterpret_cast<uint4 *>(sops);
        for (int i = threadIdx.x; i < n_ops * static_cast<int>(sizeof(DevOp) / 16); i += NTHREADS)
            odst[i] = osrc[i];
        if (threadIdx.x < kTileBuffers) {
            // plain-layout passes: one arrival (with the byte count of the bulk copies); swizzled
            // passes: every load thread's cp.async group + the publisher of the tile's facts
            mbar_init(&full[threadIdx.x], BULK ? 1 : kLoadThreads + 1);
            // direct-store passes: every worker warp releases the buffer after its last gather
            mbar_init(&empty[threadIdx.x], (pp.hdr.fused_store == 1 && !gsync) ? GT / 32 : 1);
        }
    }
    __syncthreads();

    const int low = hdr.low_bits;
    for (int r = threadIdx.x; r < (1 << (B - low)); r += NTHREADS) {
        uint64_t off = 0;
        for (int j = low; j < B; j++)
            if ((r >> (j - low)) & 1)
                off |= uint64_t(1) << hdr.tile_bits[j];
        rowoff[r] = off;
    }
    for (int t = threadIdx.x; t < GT; t += NTHREADS) {
        uint64_t sf = 0;
#pragma unroll
        for (int c = 0; c < NF; c++)
            sf ^= (uint64_t(0) - ((static_cast<uint64_t>(t) >> c) & 1u)) & hdr.store_free[c];
        sfree[t] = sf;
        for (int rd = 0; rd < kAccRounds; rd++) {
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < NF; c++)
                acc ^= (0u - ((static_cast<uint32_t>(t) >> c) & 1u)) & hdr.round_col[rd][c];
            acc_tab[rd * GT + t] = acc;
        }
    }
    const int n_rounds = pp.hdr.n_rounds;
    const uint32_t lowmask = (1u << low) - 1u;
    constexpr bool plain = BULK;
    __syncthreads();

    // tiles of this CTA: k = 0 .. n_mine-1  <->  global tile blockIdx.x + k * gridDim.x
    const uint32_t n_mine = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto tile_base_of = [&](uint32_t k) { // deposit the tile id into the non-tile index bits
        const uint32_t t = blockIdx.x + k * gridDim.x;
        uint64_t tb = 0;
        const int n_seg = pp.hdr.n_seg;
#pragma unroll 1
        for (int j = 0; j < n_seg; j++)
            tb |= static_cast<uint64_t>(t & pp.hdr.seg_mask[j]) << pp.hdr.seg_shift[j];
        return tb;
    };

    long long t_start = 0;
    if constexpr (PROF)
        t_start = clock64();
    if (threadIdx.x >= NG * GT) {
        // ---- load warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n");
        const int ptid = threadIdx.x - NG * GT;
        if constexpr (BULK) {
            // Bulk copies are issued by the first load warp alone. The other load warps must not stay
            // around as idle observers of the `empty` barriers: a warp that takes no part in the
            // hand-offs can fall two phases behind, and a parity wait cannot tell phase q from q + 2.
            if (ptid >= 32)
                return;
        }
        // load warps: HBM -> shared (swizzled slots), running ahead of the workers
        long long p_wait = 0, p_t0 = 0, p_t1 = 0, p_issue = 0;
        // conditional address toggle `ptid` of the pass, kept in registers by the first load warp
        uint64_t cx_gcm = 0, cx_gcv = ~uint64_t(0);
        uint32_t cx_vec = 0, cx_round = 0;
        if (ptid < hdr.n_cx) {
            cx_gcm = hdr.cx[ptid].gcm;
            cx_gcv = hdr.cx[ptid].gcv;
            cx_vec = hdr.cx[ptid].vec;
            cx_round = hdr.cx[ptid].round;
        }
        for (uint32_t k = 0; k < n_mine; k++) {
            const int bi = k % kTileBuffers;
            if constexpr (PROF)
                p_t0 = clock64();
            if (k >= kTileBuffers)
                mbar_wait_warp<400>(&empty[bi], ((k / kTileBuffers) - 1) & 1u);
            if constexpr (PROF) {
                p_t1 = clock64();
                p_wait += p_t1 - p_t0;
            }
            const uint64_t tb = tile_base_of(k);
            if (ptid < 32) {
                // first load warp: the tile's facts. Lane c owns conditional toggle c (registers),
                // one xor-reduction per round gives the toggles visible from that round on.
                const bool fire = ((tb | rank_bits) & cx_gcm) == cx_gcv;
                for (int r = 0; r <= n_rounds; r++) {
                    const uint32_t x =
                        __reduce_xor_sync(0xffffffffu, (fire && cx_round <= static_cast<uint32_t>(r)) ? cx_vec : 0u);
                    if (ptid == 0)
                        tinfo[bi].xoff[r] = x;
                }
                // fused store: the conditional toggles of the global index that fire for this tile
                const uint64_t sx = (fire && ptid < hdr.n_cx) ? hdr.store_cx[ptid] : uint64_t(0);
                const uint32_t sx_lo = __reduce_xor_sync(0xffffffffu, static_cast<uint32_t>(sx));
                const uint32_t sx_hi = __reduce_xor_sync(0xffffffffu, static_cast<uint32_t>(sx >> 32));
                if (ptid == 0) {
                    tinfo[bi].tb = tb;
                    tinfo[bi].store_base = tb ^ (static_cast<uint64_t>(sx_hi) << 32 | sx_lo);
                    if (plain)
                        mbar_arrive_expect_tx(&full[bi], static_cast<uint32_t>(TILE * sizeof(amp_t)));
                    else
                        mbar_arrive(&full[bi]); // release: the stores above are visible to whoever acquires
                }
            }
            if constexpr (plain) {
                // plain layout: a tile is 2^(B-low) runs of 2^low amplitudes, each contiguous in HBM and
                // in shared memory -- one bulk async copy per run, no per-amplitude address arithmetic
                amp_t *buf = tiles + bi * TILE;
                const int run = pp.hdr.bulk_run_bits; // leading tile bits that are index bits 0..run-1
                const uint32_t run_bytes = static_cast<uint32_t>(sizeof(amp_t)) << run;
                for (int r = ptid; r < (TILE >> run); r += 32)
                    bulk_load(buf + (r << run), state + (tb | rowoff[r << (run - low)]), run_bytes, &full[bi]);
                if constexpr (PROF)
                    p_issue += clock64() - p_t1;
            } else {
            // element i = e * 128 + ptid: the thread part of both addresses is loop-invariant, the
            // e part is a compile-time slot offset and a uniform (constant-bank) index offset
            // The tile is copied in 16-byte units (one complex128, two complex64 that the swizzle keeps
            // together): unit q = e * 128 + ptid holds tile-local index q << SH and lands in unit slot
            // phys<B - SH, SW - SH>(q) = (phys(ptid) ^ f_e) + e * 128 with f_e the swizzle fold of e * 128
            // (a compile-time constant below 8). The copies are issued grouped by f_e, so that each is one
            // LDGSTS at an immediate offset of a per-group base address.
            constexpr int BU = B - SH, FW = SW - SH; // log2(units per tile), swizzled unit bits
            constexpr uint32_t UPT = (1u << BU) / kLoadThreads;
            const uint32_t dbase = smem_u32(smem_raw) + static_cast<uint32_t>(bi) * ((1u << BU) * 16u);
            const uint32_t slotB = phys<BU, FW>(static_cast<uint32_t>(ptid)) * 16u;
            const uint32_t li = static_cast<uint32_t>(ptid) << SH; // tile-local index of the thread's part
            const char *src = reinterpret_cast<const char *>(state + (tb | rowoff[li >> low] | (li & lowmask)));
#pragma unroll
            for (int f = 0; f < (1 << FW); f++) {
                const uint32_t d = dbase + (slotB ^ (static_cast<uint32_t>(f) * 16u));
#pragma unroll
                for (uint32_t e = 0; e < UPT; e++)
                    if ((phys<BU, FW>(e * kLoadThreads) & ((1u << FW) - 1u)) == static_cast<uint32_t>(f))
                        cp_async_to<16>(d + e * kLoadThreads * 16u, src + pp.hdr.load_off[e]);
            }
            cp_async_arrive(&full[bi]);
            if constexpr (PROF)
                p_issue += clock64() - p_t1;
            }
        }
        if constexpr (PROF) {
            if (ptid == 0) {
                const long long life = clock64() - t_start;
                atomicAdd(&g_tile_prof[11], static_cast<unsigned long long>(p_issue));
                atomicAdd(&g_tile_prof[2], static_cast<unsigned long long>(p_wait));
                atomicAdd(&g_tile_prof[3], static_cast<unsigned long long>(life - p_wait));
                atomicAdd(&g_tile_prof[5], static_cast<unsigned long long>(life));
            }
        }
        return;
    }

    // ---- worker groups: tile k is computed by group k & 1
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    const int grp = threadIdx.x / GT;
    const int tid = threadIdx.x % GT;
    long long w_wait = 0, w_t0 = 0, w_tiles = 0, w_last = 0, w_round = 0, w_r0 = 0, w_pro = 0, w_p0 = 0;
    for (uint32_t k = grp; k < n_mine; k += NG) {
        if constexpr (PROF)
            w_p0 = clock64();
        const int bi = k % kTileBuffers;
        amp_t *tile = tiles + bi * TILE;
        const uint32_t *xoff = tinfo[bi].xoff;
        if constexpr (PROF) {
            w_t0 = clock64();
            w_pro += w_t0 - w_p0;
        }
        mbar_wait_warp<64>(&full[bi], (k / kTileBuffers) & 1u);
        if (gsync)
            group_sync(1 + grp, GT);
        const uint64_t tb = tinfo[bi].tb;
        const uint64_t tbr = tb | rank_bits;
        if constexpr (PROF) {
            w_wait += clock64() - w_t0;
            w_tiles++;
        }

#pragma unroll 1
        for (int rd = 0; rd < n_rounds; rd++) {
            // this thread's register group: logical base index (high half), storage slot (low half)
            uint32_t acc = 0;
            if (rd < kAccRounds) {
                acc = acc_tab[rd * GT + tid];
            } else {
#pragma unroll
                for (int c = 0; c < NF; c++)
                    acc ^= (0u - ((static_cast<uint32_t>(tid) >> c) & 1u)) & hdr.round_col[rd][c];
            }
            const uint32_t base = acc >> 16;
            // Byte offsets from the start of the tile buffers: the buffer's own offset is a multiple of the
            // tile size, so it joins the XOR and every gather / scatter address is one LOP3 away from pb.
            constexpr uint32_t AB = static_cast<uint32_t>(sizeof(amp_t));
            const uint32_t pb = (((acc & 0xffffu) ^ xoff[rd]) * AB) ^ (static_cast<uint32_t>(bi) * (TILE * AB));
            uint32_t poff[R];
#pragma unroll
            for (int s = 0; s < R; s++)
                poff[s] = static_cast<uint32_t>(hdr.round_poff[rd][s]) * AB;
            const int o_begin = pp.hdr.round_begin[rd], o_end = pp.hdr.round_begin[rd + 1];
            if constexpr (PROF)
                w_r0 = clock64();
            amp_t a[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                uint32_t x = pb;
#pragma unroll
                for (int c = 0; c < R; c++)
                    if (s & (1 << c))
                        x ^= poff[c];
                a[s] = *reinterpret_cast<const amp_t *>(smem_raw + x);
            }
            // a zero the compiler cannot see through: the 2^R scatter addresses are recomputed
            // after the arithmetic instead of being kept alive -- and spilled -- across the round
            const uint32_t opaque_zero = pp.hdr.pad_[0];
            const int kind = pp.hdr.round_kind[rd];
            // fused store: the last round's registers go straight to HBM; the tile buffer is free
            // as soon as every thread of the group has gathered from it
            const int fused = rd == n_rounds - 1 ? pp.hdr.fused_store : 0;
            uint64_t sbase = 0; // global index of register slot 0 in the fused store (read before the release)
            if (fused == 1)
                sbase = tinfo[bi].store_base ^ sfree[tid];
            if (fused == 1 && gsync) {
                if constexpr (plain)
                    fence_proxy_async();
                group_sync(1 + grp, GT);
                if (tid == 0)
                    mbar_arrive(&empty[bi]);
            } else if (fused == 1) {
                // no group barrier: the buffer is handed back once all warps of the group arrived,
                // and a warp whose stores are out moves on to its next tile on its own
                if constexpr (plain)
                    fence_proxy_async();
                __syncwarp();
                if ((tid & 31) == 0)
                    mbar_arrive(&empty[bi]);
            }
            if (kind == 0) {
                if constexpr (INTERP) {
#pragma unroll 1
                    for (int oi = o_begin; oi < o_end; oi++)
                        run_op<R, NS, amp_t, real>(a, sops[oi], tbr, base);
                    finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                } else {
                    __trap(); // the launcher picked a variant without the interpreter
                }
            } else if (kind >= 8) {
                if constexpr (!FACT)
                    __trap(); // the launcher picked a variant without the factored bodies
                // factored dense round; the first four rounds of a pass use compile-time table offsets (LDCU),
                // later ones index the table at run time (LDC). Measured: giving rounds 4 and 5 fixed
                // offsets too makes ptxas' code for the whole loop 10 % slower, so they stay generic.
                bool done = false;
                if constexpr (FACT)
                switch (rd) {
#define B2_RD(RD)                                                                               \
    case RD:                                                                                    \
        done = factored_round_fixed<RD, R, NS, amp_t, real>(                                    \
            kind, a, pp, fused, state, smem_raw, pb ^ opaque_zero, poff, sbase);                \
        break;
                    B2_RD(0)
                    B2_RD(1)
                    B2_RD(2)
                    B2_RD(3)
#undef B2_RD
                default:
                    break;
                }
                if (FACT && !done) {
                    switch (kind) {
#define B2_FACT(G)                                                                              \
    case 8 + G:                                                                                 \
        if constexpr (FACT && G <= R) {                                                         \
            dense_factored<G, R, NS, amp_t, real>(a, pp.dense[rd]);                             \
            finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr,      \
                                sbase);                                                         \
        }                                                                                       \
        break;
                        B2_FACT(2)
                        B2_FACT(3)
                        B2_FACT(4)
                        B2_FACT(5)
#undef B2_FACT
                    default:
                        break;
                    }
                }
            } else {
                // dense round: gate k acts on register slot k; each case is straight-line code from
                // the gathered registers to the scatter, so ptxas renames freely (no moves)
                switch (kind) {
                case 1:
                    dense_round<1, R, NS, amp_t, real>(a, pp.ops + o_begin);
                    finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                    break;
                default:
                    if constexpr (DENSEK) {
                        switch (kind) {
                        case 2:
                            dense_round<2, R, NS, amp_t, real>(a, pp.ops + o_begin);
                            finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                            break;
                        case 3:
                            dense_round<3, R, NS, amp_t, real>(a, pp.ops + o_begin);
                            finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                            break;
                        case 4:
                            dense_round<4, R, NS, amp_t, real>(a, pp.ops + o_begin);
                            finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                            break;
                        default:
                            dense_round<5, R, NS, amp_t, real>(a, pp.ops + o_begin);
                            finish_round<R, NS>(fused, state, smem_raw, a, pb ^ opaque_zero, poff, pp.hdr, sbase);
                            break;
                        }
                    } else {
                        __trap(); // the launcher picked a variant without these bodies
                    }
                    break;
                }
            }
            if (!fused)
                group_sync(1 + grp, GT);
            if constexpr (PROF) {
                if (fused)
                    w_last += clock64() - w_r0;
                else
                    w_round += clock64() - w_r0;
            }
        }
        if (pp.hdr.fused_store && n_rounds > 0)
            continue; // stored from registers, buffer already released

        // ---- shared -> HBM through the final address map
        {
            constexpr int NT = NF; // log2(GT)
            uint32_t sl = xoff[n_rounds];
#pragma unroll
            for (int c = 0; c < NT; c++)
                sl ^= (0u - ((static_cast<uint32_t>(tid) >> c) & 1u)) & hdr.final_col[c];
            uint32_t ecol[B - NT];
#pragma unroll
            for (int c = 0; c < B - NT; c++)
                ecol[c] = hdr.final_col[NT + c];
#pragma unroll
            for (int e = 0; e < EPT; e++) {
                const uint32_t i = e * GT + tid;
                uint32_t x = sl;
#pragma unroll
                for (int c = 0; c < B - NT; c++)
                    if (e & (1 << c))
                        x ^= ecol[c];
                state[tb | rowoff[i >> low] | (i & lowmask)] = tile[x];
            }
        }
        if constexpr (plain)
            fence_proxy_async();
        group_sync(1 + grp, GT); // every thread of the group is done with this buffer (and xoff)
        if (tid == 0)
            mbar_arrive(&empty[bi]);
    }
    if constexpr (PROF) {
        if (tid == 0) {
            const long long life = clock64() - t_start;
            atomicAdd(&g_tile_prof[0], static_cast<unsigned long long>(w_wait));
            atomicAdd(&g_tile_prof[1], static_cast<unsigned long long>(life - w_wait));
            atomicAdd(&g_tile_prof[4], static_cast<unsigned long long>(w_tiles));
            atomicAdd(&g_tile_prof[6], static_cast<unsigned long long>(w_last));
            atomicAdd(&g_tile_prof[7], static_cast<unsigned long long>(w_round));
            atomicAdd(&g_tile_prof[10], static_cast<unsigned long long>(w_pro));
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------
namespace {
template <typename real, int B, int NB, int GT> constexpr size_t tile_smem_bytes() {
    return NB * sizeof(typename AmpT<real>::type) * (size_t(1) << B) +
           sizeof(DevOp) * kMaxOpsPerPass + sizeof(uint64_t) * (size_t(1) << (B - kMinLow)) +
           sizeof(uint64_t) * GT + sizeof(uint32_t) * kAccRounds * GT;
}
int sm_count() { return sm_count_current_device(); }
int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
bool tile_prof() {
    static const bool v = env_int("B2SV_TILE_PROF", 0) != 0;
    return v;
}
template <typename real, int B, int R, int GT, int NG, int NB, bool FACT, bool INTERP, bool DENSEK,
          bool PROF, bool BULK>
void launch_variant(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                    cudaStream_t stream, int max_ctas) {
    using amp_t = typename AmpT<real>::type;
    auto kern = tile_exec_kernel<real, B, R, GT, NG, NB, FACT, INTERP, DENSEK, PROF, BULK>;
    constexpr size_t smem = tile_smem_bytes<real, B, NB, GT>();
    static_assert(smem + 6 * 1024 <= 227 * 1024, "dynamic + static shared memory of the tile kernel");
    static uint64_t configured = 0; // one bit per device: the attribute is per device
    if (first_use_on_device(configured))
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    const uint32_t n_tiles = 1u << (n_eff - B);
    // one persistent CTA per SM; with fewer than 2 tiles per SM spread them one per CTA
    // (max_ctas > 0: leave SMs to a kernel running beside this one, e.g. an exchange between shards)
    const unsigned grid = std::min<uint32_t>(
        n_tiles, static_cast<uint32_t>(max_ctas > 0 ? std::min(max_ctas, sm_count()) : sm_count()));
    // the descriptor is copied into the launch's parameter buffer by the runtime at this call
    constexpr int nthreads = NG * GT + kLoadThreads;
    static const uint32_t flags = static_cast<uint32_t>(env_int("B2SV_TILE_FLAGS", 0));
    kern<<<grid, nthreads, smem, stream>>>(static_cast<amp_t *>(state), pp, rank_bits, n_tiles, flags);
    CUDA_CHECK(cudaGetLastError());
}

// Picks the leanest kernel variant that covers the rounds of the pass.
template <typename real, int B, int R, bool BULK>
void launch_tile_pass_v(void *state, const PassParams &pp, int n_eff, uint64_t rank_bits,
                        cudaStream_t stream, int max_ctas) {
    B2_ABORT_IF(pp.hdr.low_bits < kMinLow, "tile executor: fewer contiguous low bits than the row table is sized for");
    bool fact = false, interp = false, densek = false;
    for (int rd = 0; rd < pp.hdr.n_rounds; rd++) {
        const int kind = pp.hdr.round_kind[rd];
        fact |= kind >= 8;
        interp |= kind == 0;
        densek |= kind >= 2 && kind < 8;
    }
    static const bool debug = env_int("B2SV_DEBUG_PASSES", 0) != 0;
    if (debug) { // one line per launch: variant flags, layout, round kinds (then synchronise to localise faults)
        fprintf(stderr, "b2sv pass: fact=%d interp=%d densek=%d bulk=%d run_bits=%d fused=%d n_cx=%d n_ops=%d kinds=", fact,
                interp, densek, int(BULK), int(pp.hdr.bulk_run_bits), int(pp.hdr.fused_store), pp.hdr.n_cx, pp.hdr.n_ops);
        for (int rd = 0; rd < pp.hdr.n_rounds; rd++)
            fprintf(stderr, "%d,", int(pp.hdr.round_kind[rd]));
        fprintf(stderr, " tile_bits=");
        for (int j = 0; j < B; j++)
            fprintf(stderr, "%d,", int(pp.hdr.tile_bits[j]));
        fprintf(stderr, "\n");
        cudaError_t e0 = cudaStreamSynchronize(stream);
        if (e0 != cudaSuccess)
            fprintf(stderr, "b2sv pass: error BEFORE this launch: %s\n", cudaGetErrorString(e0));
    }
    if (tile_prof())
        launch_variant<real, B, R, 256, 2, 3, true, true, true, true, false>(state, pp, n_eff, rank_bits, stream, max_ctas);
    else if (fact && !interp && !densek) // layered circuits: factored rounds (+ single gates)
        launch_variant<real, B, R, 256, 2, 3, true, false, false, false, BULK>(state, pp, n_eff, rank_bits, stream, max_ctas);
    else if (!fact && !interp)           // unfactored dense rounds, single gates, permutation-only passes
        launch_variant<real, B, R, 256, 2, 3, false, false, true, false, BULK>(state, pp, n_eff, rank_bits, stream, max_ctas);
    else if (!fact)                      // controlled / diagonal ops through the interpreter
        launch_variant<real, B, R, 256, 2, 3, false, true, true, false, BULK>(state, pp, n_eff, rank_bits, stream, max_ctas);
    else
        launch_variant<real, B, R, 256, 2, 3, true, true, true, false, BULK>(state, pp, n_eff, rank_bits, stream, max_ctas);
}
} // namespace

} // namespace b2sv
