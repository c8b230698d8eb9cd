// b2sv: the tile executor -- ONE kernel applies every gate of a fused pass in a single HBM sweep.
//
// Replaces the reference's one-gate-per-sweep functor dispatch
// (reference StateVectorKokkos.hpp:807-824 applyGateFunctor -> GateFunctors.hpp, one
// Kokkos::parallel_for over 2^(n-k) per gate).
//
// Data flow per CTA (one tile of 2^B amplitudes, B = 12 for complex128 -> 64 KiB of shared memory):
//   HBM --coalesced 16 B/lane loads, 2^low-amplitude contiguous runs--> shared memory (XOR-swizzled)
//   for each round: every thread gathers 2^R amplitudes whose indices differ only in the round's R
//     "register bits", applies all ops of the round in registers, scatters back in place
//   shared memory --> HBM (same addresses: the update is in place)
// HBM traffic is exactly one read + one write of the state per pass, whatever the number of gates.
//
// Shared-memory layout: amplitude i of the tile lives at slot phys(i) = i ^ fold(i) where fold XORs
// the higher SW-bit groups of i into its low SW bits (SW = log2(128 B / sizeof(amp))). phys is
// GF(2)-linear, so phys(base | off) = phys(base) ^ phys(off), and any 8 (c128) / 16 (c64) consecutive
// lanes hit distinct 16 B / 8 B bank groups for every choice of register bits.
#include "schedule.hpp"

#include <cuda_runtime.h>

namespace b2sv {

template <typename real> struct AmpT;
template <> struct AmpT<double> {
    using type = double2;
};
template <> struct AmpT<float> {
    using type = float2;
};

template <int B, int SW> __device__ __forceinline__ uint32_t phys(uint32_t i) {
    uint32_t f = 0;
#pragma unroll
    for (int s = SW; s < B; s += SW)
        f ^= (i >> s);
    return i ^ (f & ((1u << SW) - 1u));
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

// ---- the kernel ---------------------------------------------------------------------------------
template <typename real, int B, int R, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    tile_exec_kernel(typename AmpT<real>::type *__restrict__ state,
                     const unsigned char *__restrict__ blob, uint64_t rank_bits) {
    using amp_t = typename AmpT<real>::type;
    constexpr int SW = (sizeof(amp_t) == 16) ? 3 : 4;
    constexpr int NS = 1 << R;
    constexpr int TILE = 1 << B;
    constexpr int NF = B - R; // non-register tile bits = thread-id bits
    static_assert((1 << NF) == THREADS, "one register group per thread");
    static_assert(NF <= kMaxFreeBits, "too many thread-id bits");
    constexpr int EPT = TILE / THREADS; // amplitudes per thread in the load / store phases (= NS)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    amp_t *tile = reinterpret_cast<amp_t *>(smem_raw);
    DevOp *sops = reinterpret_cast<DevOp *>(smem_raw + sizeof(amp_t) * TILE);
    uint64_t *rowoff = reinterpret_cast<uint64_t *>(sops + kMaxOpsPerPass);
    __shared__ DevPassHeader hdr;
    __shared__ uint32_t xoff[kMaxRounds + 1];

    const int tid = threadIdx.x;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(&hdr);
        for (int i = tid; i < static_cast<int>(sizeof(DevPassHeader) / 16); i += THREADS)
            dst[i] = src[i];
        const int n_ops = reinterpret_cast<const DevPassHeader *>(blob)->n_ops;
        const uint4 *osrc = reinterpret_cast<const uint4 *>(blob + sizeof(DevPassHeader));
        uint4 *odst = reinterpret_cast<uint4 *>(sops);
        for (int i = tid; i < n_ops * static_cast<int>(sizeof(DevOp) / 16); i += THREADS)
            odst[i] = osrc[i];
    }
    __syncthreads();

    const int low = hdr.low_bits;
    // tile id -> index with the tile bits cleared (deposit into the non-tile positions)
    uint64_t tb = blockIdx.x;
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        const int p = hdr.tile_bits[j];
        tb = ((tb >> p) << (p + 1)) | (tb & ((uint64_t(1) << p) - 1));
    }
    for (int r = tid; r < (1 << (B - low)); r += THREADS) {
        uint64_t off = 0;
        for (int j = low; j < B; j++)
            if ((r >> (j - low)) & 1)
                off |= uint64_t(1) << hdr.tile_bits[j];
        rowoff[r] = off;
    }
    const uint64_t tbr = tb | rank_bits;
    const int n_rounds = hdr.n_rounds;
    if (tid <= n_rounds) { // CTA-uniform address toggles visible from round `tid` on
        uint32_t x = 0;
        for (int k = 0; k < hdr.n_cx; k++)
            if (hdr.cx[k].round <= tid && (tbr & hdr.cx[k].gcm) == hdr.cx[k].gcv)
                x ^= hdr.cx[k].vec;
        xoff[tid] = x;
    }
    __syncthreads();

    // ---- HBM -> shared (identity address map)
    const uint32_t lowmask = (1u << low) - 1u;
    {
        amp_t v[EPT];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const uint32_t i = e * THREADS + tid;
            v[e] = state[tb | rowoff[i >> low] | (i & lowmask)];
        }
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const uint32_t i = e * THREADS + tid;
            tile[phys<B, SW>(i)] = v[e];
        }
    }
    __syncthreads();

#pragma unroll 1
    for (int rd = 0; rd < n_rounds; rd++) {
        // this thread's register group: logical base index (high half) and storage slot (low half)
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < NF; k++)
            acc ^= (0u - ((static_cast<uint32_t>(tid) >> k) & 1u)) & hdr.round_col[rd][k];
        const uint32_t base = acc >> 16;
        const uint32_t pb = (acc & 0xffffu) ^ xoff[rd];
        uint32_t poff[R];
#pragma unroll
        for (int s = 0; s < R; s++)
            poff[s] = hdr.round_poff[rd][s];
        const int o_begin = hdr.round_begin[rd], o_end = hdr.round_begin[rd + 1];
        amp_t a[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            uint32_t x = pb;
#pragma unroll
            for (int k = 0; k < R; k++)
                if (s & (1 << k))
                    x ^= poff[k];
            a[s] = tile[x];
        }
#pragma unroll 1
        for (int oi = o_begin; oi < o_end; oi++)
            run_op<R, NS, amp_t, real>(a, sops[oi], tbr, base);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            uint32_t x = pb;
#pragma unroll
            for (int k = 0; k < R; k++)
                if (s & (1 << k))
                    x ^= poff[k];
            tile[x] = a[s];
        }
        __syncthreads();
    }

    // ---- shared -> HBM through the final address map
    {
        constexpr int NT = NF; // log2(THREADS)
        uint32_t sl = xoff[n_rounds];
#pragma unroll
        for (int k = 0; k < NT; k++)
            sl ^= (0u - ((static_cast<uint32_t>(tid) >> k) & 1u)) & hdr.final_col[k];
        uint32_t ecol[B - NT];
#pragma unroll
        for (int k = 0; k < B - NT; k++)
            ecol[k] = hdr.final_col[NT + k];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const uint32_t i = e * THREADS + tid;
            uint32_t x = sl;
#pragma unroll
            for (int k = 0; k < B - NT; k++)
                if (e & (1 << k))
                    x ^= ecol[k];
            state[tb | rowoff[i >> low] | (i & lowmask)] = tile[x];
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------
namespace {
constexpr int kMinLow = 4;
template <typename real, int B> constexpr size_t tile_smem_bytes() {
    return sizeof(typename AmpT<real>::type) * (size_t(1) << B) + sizeof(DevOp) * kMaxOpsPerPass +
           sizeof(uint64_t) * (size_t(1) << (B - kMinLow));
}
} // namespace

// Tile geometry per dtype: complex128 -> 2^12 amps (64 KiB), complex64 -> 2^13 amps (64 KiB).
void tile_config(int dtype, int *B, int *R) {
    if (dtype == 1) {
        *B = 12;
        *R = 4;
    } else {
        *B = 13;
        *R = 4;
    }
}

void launch_tile_pass(int dtype, void *state, const unsigned char *dev_blob, int n_eff,
                      uint64_t rank_bits, cudaStream_t stream) {
    if (dtype == 1) {
        constexpr int B = 12;
        auto kern = tile_exec_kernel<double, B, 4, 256, 2>;
        constexpr size_t smem = tile_smem_bytes<double, B>();
        static bool configured = false;
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
            configured = true;
        }
        const unsigned grid = 1u << (n_eff - B);
        kern<<<grid, 256, smem, stream>>>(static_cast<double2 *>(state), dev_blob, rank_bits);
    } else {
        constexpr int B = 13;
        auto kern = tile_exec_kernel<float, B, 4, 512, 2>;
        constexpr size_t smem = tile_smem_bytes<float, B>();
        static bool configured = false;
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
            configured = true;
        }
        const unsigned grid = 1u << (n_eff - B);
        kern<<<grid, 512, smem, stream>>>(static_cast<float2 *>(state), dev_blob, rank_bits);
    }
    CUDA_CHECK(cudaGetLastError());
}

} // namespace b2sv
