// b2sv: stand-alone CUDA kernels -- state initialisation, generic k-qubit matrices, reductions
// (expectation values, inner products), Pauli-sum application, CSR SpMV, probabilities, sampling.
//
// Reference counterparts (one Kokkos functor each):
//   InitView / setBasisStateFunctor / setStateVectorFunctor   StateVectorKokkos.hpp:46-90
//   multiQubitOpFunctor                                       GateFunctors.hpp:195-300
//   getExpectationValue*Functor                               ExpValFunctors.hpp:13-280
//   getReal/ImagOfComplexInnerProductFunctor, axpy, SparseMV  LinearAlgebraKokkos.hpp:30-236
//   getProbFunctor / getSubProbFunctor / getCDFFunctor / Sampler   MeasuresFunctors.hpp:15-210
// All reductions accumulate in double (also for complex64), each block writes its partial sums
// and a single-block finalize kernel adds them in a fixed order: results are deterministic.
#include "kernels.cuh"
#include <algorithm>
#include "common.hpp"

namespace b2sv {
namespace {

template <typename real> struct AmpT;
template <> struct AmpT<double> {
    using type = double2;
};
template <> struct AmpT<float> {
    using type = float2;
};

__device__ __forceinline__ uint64_t insert_zero(uint64_t k, int p) {
    return ((k >> p) << (p + 1)) | (k & ((uint64_t(1) << p) - 1));
}

// ---- block reduction of NV doubles, result written to partials[blockIdx.x * NV + v]
template <int NV> __device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partials) {
    __shared__ double red[NV][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        double x = v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0)
            red[j][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            double x = lane < nw ? red[j][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0)
                partials[blockIdx.x * NV + j] = x;
        }
    }
}

__global__ void k_finalize(const double *__restrict__ partials, int nblocks, int nv,
                           double *__restrict__ out) {
    // one block; thread t sums blocks t, t+256, ... in a fixed order
    __shared__ double red[kReduceThreads];
    for (int j = 0; j < nv; j++) {
        double x = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
            x += partials[b * nv + j];
        red[threadIdx.x] = x;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s)
                red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0)
            out[j] = red[0];
        __syncthreads();
    }
}

// one block per value (for reductions with many values): block j sums column j in a fixed order
__global__ void k_finalize_wide(const double *__restrict__ partials, int nblocks, int nv,
                                double *__restrict__ out) {
    __shared__ double red[kReduceThreads];
    const int j = blockIdx.x;
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        x += partials[b * nv + j];
    red[threadIdx.x] = x;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        out[j] = red[0];
}

// ---- state management ---------------------------------------------------------------------------
template <typename amp_t>
__global__ void k_set_basis(amp_t *__restrict__ state, uint64_t len, uint64_t index) {
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        amp_t v;
        v.x = (i == index) ? 1 : 0;
        v.y = 0;
        state[i] = v;
    }
}
template <typename amp_t>
__global__ void k_scatter(amp_t *__restrict__ state, const uint64_t *__restrict__ idx,
                          const double2 *__restrict__ val, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        amp_t v;
        v.x = val[i].x;
        v.y = val[i].y;
        state[idx[i]] = v;
    }
}
// State preparation on a subset of wires, index table built on the device (the reference's Python
// layer builds it with itertools.product and ships 2^k indices, lightning_kokkos.py:317-327):
// state[sum_j bit_j(v) << pos[j]] = val[v], v < 2^k, pos[j] = index bit of the j-th wire (wire 0 of
// the list = MSB of v). Only the entries of this shard (index >> n_local == rank) are written.
struct WirePos {
    int k;
    int pos[64];
};
template <typename amp_t>
__global__ void k_scatter_wires(amp_t *__restrict__ state, const double2 *__restrict__ val, uint64_t count,
                                WirePos wp, int n_local, uint64_t rank) {
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t v = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; v < count; v += stride) {
        uint64_t idx = 0;
        for (int j = 0; j < wp.k; j++)
            idx |= ((v >> (wp.k - 1 - j)) & 1ull) << wp.pos[j];
        if ((n_local >= 64 ? 0 : (idx >> n_local)) == rank) {
            amp_t a;
            a.x = val[v].x;
            a.y = val[v].y;
            state[idx & ((uint64_t(1) << n_local) - 1)] = a;
        }
    }
}
// out[k] = state[idx[k]] when the index belongs to this shard (idx >> n_local == rank), else 0
template <typename amp_t>
__global__ void k_gather(const amp_t *__restrict__ state, const uint64_t *__restrict__ idx, size_t n,
                         int n_local, uint64_t rank, double2 *__restrict__ out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t g = idx[i];
        double2 v = make_double2(0.0, 0.0);
        if ((n_local >= 64 ? 0 : (g >> n_local)) == rank) {
            const amp_t a = state[g & ((uint64_t(1) << n_local) - 1)];
            v = make_double2(a.x, a.y);
        }
        out[i] = v;
    }
}
template <typename amp_t, typename real>
__global__ void k_axpy(real ar, real ai, const amp_t *__restrict__ x, amp_t *__restrict__ y,
                       uint64_t len) {
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t a = x[i];
        amp_t b = y[i];
        b.x += ar * a.x - ai * a.y;
        b.y += ar * a.y + ai * a.x;
        y[i] = b;
    }
}

// ---- generic k-qubit matrix ---------------------------------------------------------------------
struct MatKParams {
    int k;
    int bits[10];   // bits[0] = MSB of the local index
    int sorted[10]; // ascending
};
// One CTA works on G = max(1, blockDim/dim) groups at a time: gather 2^k amps per group to shared
// memory, each thread produces one output row (looping when dim > blockDim), scatter back.
template <typename amp_t>
__global__ void k_matk(amp_t *__restrict__ state, const double2 *__restrict__ mat, MatKParams p,
                       uint64_t ngroups) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *v = reinterpret_cast<double2 *>(smem_raw);
    const int dim = 1 << p.k;
    const int G = blockDim.x >= dim ? blockDim.x / dim : 1;
    const uint64_t niter = (ngroups + G - 1) / G;
    for (uint64_t it = blockIdx.x; it < niter; it += gridDim.x) {
        // gather
        for (int e = threadIdx.x; e < G * dim; e += blockDim.x) {
            const int gs = e / dim, c = e % dim;
            const uint64_t grp = it * G + gs;
            if (grp < ngroups) {
                uint64_t base = grp;
                for (int j = 0; j < p.k; j++)
                    base = insert_zero(base, p.sorted[j]);
                uint64_t off = 0;
                for (int j = 0; j < p.k; j++)
                    if ((c >> (p.k - 1 - j)) & 1)
                        off |= uint64_t(1) << p.bits[j];
                const amp_t a = state[base | off];
                v[e] = make_double2(a.x, a.y);
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < G * dim; e += blockDim.x) {
            const int gs = e / dim, r = e % dim;
            const uint64_t grp = it * G + gs;
            if (grp < ngroups) {
                double sr = 0.0, si = 0.0;
                const double2 *row = mat + size_t(r) * dim;
                const double2 *vin = v + gs * dim;
                for (int c = 0; c < dim; c++) {
                    const double2 m = row[c];
                    const double2 x = vin[c];
                    sr += m.x * x.x - m.y * x.y;
                    si += m.x * x.y + m.y * x.x;
                }
                uint64_t base = grp;
                for (int j = 0; j < p.k; j++)
                    base = insert_zero(base, p.sorted[j]);
                uint64_t off = 0;
                for (int j = 0; j < p.k; j++)
                    if ((r >> (p.k - 1 - j)) & 1)
                        off |= uint64_t(1) << p.bits[j];
                amp_t o;
                o.x = sr;
                o.y = si;
                state[base | off] = o;
            }
        }
        __syncthreads();
    }
}

// ---- reductions ---------------------------------------------------------------------------------
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_norm2(const amp_t *__restrict__ s, uint64_t len, double *__restrict__ partials) {
    double acc[1] = {0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t a = s[i];
        acc[0] += double(a.x) * a.x + double(a.y) * a.y;
    }
    block_reduce_store<1>(acc, partials);
}
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_dot(const amp_t *__restrict__ x, const amp_t *__restrict__ y, uint64_t len,
          double *__restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t a = x[i], b = y[i];
        acc[0] += double(a.x) * b.x + double(a.y) * b.y; // Re conj(a) b
        acc[1] += double(a.x) * b.y - double(a.y) * b.x; // Im conj(a) b
    }
    block_reduce_store<2>(acc, partials);
}
struct M8 {
    double m[8];
};
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_expval_1q(const amp_t *__restrict__ s, uint64_t npairs, int tbit, M8 mm,
                double *__restrict__ partials) {
    double acc[1] = {0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < npairs; k += stride) {
        const uint64_t i0 = insert_zero(k, tbit), i1 = i0 | (uint64_t(1) << tbit);
        const amp_t a0 = s[i0], a1 = s[i1];
        const double v0r = a0.x, v0i = a0.y, v1r = a1.x, v1i = a1.y;
        // w = M v ; acc += Re(conj(v) . w)
        const double w0r = mm.m[0] * v0r - mm.m[1] * v0i + mm.m[2] * v1r - mm.m[3] * v1i;
        const double w0i = mm.m[0] * v0i + mm.m[1] * v0r + mm.m[2] * v1i + mm.m[3] * v1r;
        const double w1r = mm.m[4] * v0r - mm.m[5] * v0i + mm.m[6] * v1r - mm.m[7] * v1i;
        const double w1i = mm.m[4] * v0i + mm.m[5] * v0r + mm.m[6] * v1i + mm.m[7] * v1r;
        acc[0] += v0r * w0r + v0i * w0i + v1r * w1r + v1i * w1i;
    }
    block_reduce_store<1>(acc, partials);
}
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_expval_2q(const amp_t *__restrict__ s, uint64_t nquads, int bit_a, int bit_b,
                const double2 *__restrict__ m16, double *__restrict__ partials) {
    __shared__ double2 sm[16];
    if (threadIdx.x < 16)
        sm[threadIdx.x] = m16[threadIdx.x];
    __syncthreads();
    const int lo = bit_a < bit_b ? bit_a : bit_b, hi = bit_a < bit_b ? bit_b : bit_a;
    double acc[1] = {0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nquads; k += stride) {
        const uint64_t base = insert_zero(insert_zero(k, lo), hi);
        double vr[4], vi[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { // local index c = (a_bit << 1) | b_bit
            const uint64_t idx = base | (uint64_t((c >> 1) & 1) << bit_a) | (uint64_t(c & 1) << bit_b);
            const amp_t a = s[idx];
            vr[c] = a.x;
            vi[c] = a.y;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double wr = 0.0, wi = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const double2 m = sm[r * 4 + c];
                wr += m.x * vr[c] - m.y * vi[c];
                wi += m.x * vi[c] + m.y * vr[c];
            }
            acc[0] += vr[r] * wr + vi[r] * wi;
        }
    }
    block_reduce_store<1>(acc, partials);
}
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_pauli_expval(const amp_t *__restrict__ s, uint64_t len, uint64_t x, uint64_t z, double phr,
                   double phi, double *__restrict__ partials) {
    double acc[1] = {0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const uint64_t j = i ^ x;
        const amp_t a = s[i], b = s[j];
        // (P psi)_i = ph * (-1)^popc(j & z) * psi_j
        const double sg = (__popcll(j & z) & 1) ? -1.0 : 1.0;
        const double wr = sg * (phr * b.x - phi * b.y);
        const double wi = sg * (phr * b.y + phi * b.x);
        acc[0] += double(a.x) * wr + double(a.y) * wi;
    }
    block_reduce_store<1>(acc, partials);
}

// All single-qubit <Z> of a state in ONE read pass (the reference runs one reduction kernel per
// observable, MeasuresKokkos.hpp:167-271 via lightning_kokkos.py:554-559).
// A block walks chunks of 4096 amplitudes; thread t reads indices base + t + 256 j (j < 16), so
// bits 0..7 of the index are the thread's own, bits 8..11 follow j and bits >= 12 are uniform per
// chunk. Per thread: tot = sum |a|^2, four sums for bits 8..11, and for every higher bit the sum over
// the chunks that have it set. Result per block: [tot, P1(bit 0), ..., P1(bit n-1)] with
// P1(b) = sum of |a_i|^2 over indices with bit b set;  <Z_b> = tot - 2 P1(b).
constexpr int kZAllMaxBits = 40;
constexpr int kZAllVals = 1 + kZAllMaxBits;
static_assert(kZAllVals == kZAllValsHost, "kernels.cuh and kernels.cu disagree on the all-Z layout");
template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_expval_z_all(const amp_t *__restrict__ s, int n, double *__restrict__ partials) {
    constexpr int NHI = kZAllMaxBits - 12;
    double tot = 0.0, mid[4] = {0.0, 0.0, 0.0, 0.0}, hi[NHI];
#pragma unroll
    for (int b = 0; b < NHI; b++)
        hi[b] = 0.0;
    const uint64_t nchunks = uint64_t(1) << (n - 12);
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const amp_t *p = s + (c << 12) + threadIdx.x;
        double ct = 0.0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const amp_t a = p[j * 256];
            const double w = double(a.x) * a.x + double(a.y) * a.y;
            ct += w;
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (j & (1 << b))
                    mid[b] += w;
        }
        tot += ct;
#pragma unroll
        for (int b = 0; b < NHI; b++)
            if (b < n - 12 && ((c >> b) & 1ull))
                hi[b] += ct;
    }
    double v[kZAllVals];
    v[0] = tot;
#pragma unroll
    for (int b = 0; b < 8; b++)
        v[1 + b] = ((threadIdx.x >> b) & 1) ? tot : 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        v[9 + b] = mid[b];
#pragma unroll
    for (int b = 0; b < NHI; b++)
        v[13 + b] = hi[b];
    block_reduce_store<kZAllVals>(v, partials);
}

// <psi| sum_t c_t P_t |psi> for a sum of Pauli words in one kernel and without a work vector (the
// reference applies the Hamiltonian term by term into two temporaries and takes an inner product,
// ObservablesKokkos.hpp:360-373 + MeasuresKokkos.hpp:354-360). Terms are sorted by x mask on the host,
// so the partner amplitude psi[i ^ x] is re-read only when the mask changes.
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_pauli_sum_expval(const amp_t *__restrict__ s, uint64_t len, const PauliTerm *__restrict__ terms,
                       int nterms, double *__restrict__ partials) {
    constexpr int CH = 128;
    __shared__ PauliTerm st[CH];
    double acc[1] = {0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t i0 = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t iters = (len + stride - 1) / stride;
    for (uint64_t it = 0; it < iters; it++) {
        const uint64_t i = i0 + it * stride;
        amp_t a;
        a.x = 0;
        a.y = 0;
        if (i < len)
            a = s[i];
        double sr = 0.0, si = 0.0;
        for (int t0 = 0; t0 < nterms; t0 += CH) {
            const int nt = min(CH, nterms - t0);
            __syncthreads();
            if (threadIdx.x < nt)
                st[threadIdx.x] = terms[t0 + threadIdx.x];
            __syncthreads();
            if (i < len) {
                uint64_t cur_x = 0;
                amp_t b = a;
                for (int t = 0; t < nt; t++) {
                    const uint64_t j = i ^ st[t].x;
                    if (st[t].x != cur_x) {
                        cur_x = st[t].x;
                        b = s[j];
                    }
                    const double sg = (__popcll(j & st[t].z) & 1) ? -1.0 : 1.0;
                    const double cr = sg * st[t].cr, ci = sg * st[t].ci;
                    sr += cr * b.x - ci * b.y;
                    si += cr * b.y + ci * b.x;
                }
            }
        }
        acc[0] += double(a.x) * sr + double(a.y) * si; // Re conj(a_i) (H psi)_i
    }
    block_reduce_store<1>(acc, partials);
}

// <bra| P |ket> with P a Pauli word: one read pass over both vectors, nothing written.
template <typename amp_t>
__global__ void __launch_bounds__(kReduceThreads)
    k_pauli_dot(const amp_t *__restrict__ bra, const amp_t *__restrict__ ket, uint64_t len,
                uint64_t x, uint64_t z, double phr, double phi, double *__restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const uint64_t j = i ^ x;
        const amp_t a = bra[i], b = ket[j];
        // (P ket)_i = ph * (-1)^popc(j & z) * ket_j
        const double sg = (__popcll(j & z) & 1) ? -1.0 : 1.0;
        const double wr = sg * (phr * b.x - phi * b.y);
        const double wi = sg * (phr * b.y + phi * b.x);
        acc[0] += double(a.x) * wr + double(a.y) * wi; // Re conj(a) w
        acc[1] += double(a.x) * wi - double(a.y) * wr; // Im conj(a) w
    }
    block_reduce_store<2>(acc, partials);
}
// Single-qubit transition sums between two vectors, for up to NB index bits in ONE read pass
// (adjoint sweep: every trainable 1-qubit gate of a run needs <bra| M_t |ket> for some 2x2 M on
// its wire t, and all of those follow from four sums per wire).  With c_i = conj(bra_i):
//   D   = sum_i c_i ket_i                     Z_t = sum_i s_t(i) c_i ket_i      (s_t = +1 / -1 by bit t)
//   X_t = sum_i c_i ket_{i ^ 2^t}             W_t = sum_i s_t(i) c_i ket_{i ^ 2^t}
// partials per block: D (re, im), then per bit Z, X, W (re, im each) = 2 + 6 NB doubles.
struct TransBits {
    int nb;
    int pos[8];
};
template <typename amp_t, int NB>
__global__ void __launch_bounds__(kReduceThreads)
    k_transition_1q(const amp_t *__restrict__ bra, const amp_t *__restrict__ ket, uint64_t len,
                    TransBits tb, double *__restrict__ partials) {
    double acc[2 + 6 * NB];
#pragma unroll
    for (int j = 0; j < 2 + 6 * NB; j++)
        acc[j] = 0.0;
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t h = bra[i], l = ket[i];
        const double cx = double(h.x), cy = -double(h.y); // conj(bra_i)
        const double dx = cx * l.x - cy * l.y, dy = cx * l.y + cy * l.x;
        acc[0] += dx;
        acc[1] += dy;
#pragma unroll
        for (int t = 0; t < NB; t++) {
            if (t < tb.nb) {
                const double s = ((i >> tb.pos[t]) & 1ull) ? -1.0 : 1.0;
                const amp_t p = ket[i ^ (uint64_t(1) << tb.pos[t])];
                const double xx = cx * p.x - cy * p.y, xy = cx * p.y + cy * p.x;
                acc[2 + 6 * t + 0] += s * dx;
                acc[2 + 6 * t + 1] += s * dy;
                acc[2 + 6 * t + 2] += xx;
                acc[2 + 6 * t + 3] += xy;
                acc[2 + 6 * t + 4] += s * xx;
                acc[2 + 6 * t + 5] += s * xy;
            }
        }
    }
    block_reduce_store<2 + 6 * NB>(acc, partials);
}
// The same sums, tiled: a tile = the 5 lowest index bits (512-byte rows) + up to 6 chosen bits, both
// vectors' tiles staged in shared memory, so all wires of the launch (any <= kTransitionBits of the
// tile's bits) are served from ONE read of the two vectors; the untiled kernel above re-reads the
// partner amplitudes once per wire. Same partials layout as k_transition_1q.
constexpr int kTransTileBits = 11;
struct TransTile {
    int n_tile;        // number of tile bits (5 low + chosen) = kTransTileBits
    int tile_pos[12];  // ascending index bit positions of the tile
    int nb;            // wires of this launch
    int wire_tpos[8];  // position (0..n_tile-1) of each wire's bit inside the tile
    int u_pos[2];      // two tile positions that are not wires (ascending): the bits that tell a
                       // thread's four elements apart
};
__device__ __forceinline__ void cp_async_16(void *sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(sdst))),
                 "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_8(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(sdst))),
                 "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_amp(double2 *d, const double2 *s) { cp_async_16(d, s); }
__device__ __forceinline__ void cp_async_amp(float2 *d, const float2 *s) { cp_async_8(d, s); }
constexpr int kTransThreads = 512;
// One persistent CTA per SM, two shared-memory stages: the copies of tile k+1 (cp.async) are in flight
// while tile k is reduced.
// Arithmetic: thread <-> element mapping puts the tile positions of all requested wires into the
// thread-id bits (the two bits that distinguish a thread's four elements are non-wire positions), so
// s_t(e) is a per-thread constant: per element only  d += c l  (4 FMA) and, per wire,  x_t += c l_{e^t}
// (4 FMA) are accumulated; Z_t = s_t d and W_t = s_t x_t are formed once, after the loop.
template <typename amp_t>
__global__ void __launch_bounds__(kTransThreads, 1)
    k_transition_tile(const amp_t *__restrict__ bra, const amp_t *__restrict__ ket, uint64_t n_tiles,
                      TransTile tt, double *__restrict__ partials) {
    extern __shared__ __align__(16) unsigned char tsm[];
    constexpr int tile = 1 << kTransTileBits;
    constexpr int per_thread = tile / kTransThreads; // 4: two element-index bits per thread
    amp_t *stage = reinterpret_cast<amp_t *>(tsm); // [2 stages][ket tile, bra tile]
    __shared__ uint64_t rowoff[64];
    constexpr int n_rows = tile >> 5;
    if (static_cast<int>(threadIdx.x) < n_rows) {
        uint64_t off = 0;
        for (int j = 5; j < tt.n_tile; j++)
            if ((threadIdx.x >> (j - 5)) & 1)
                off |= uint64_t(1) << tt.tile_pos[j];
        rowoff[threadIdx.x] = off;
    }
    // element index of this thread's u-th element: thread-id bits deposited around the two u positions
    const int ua = tt.u_pos[0], ub = tt.u_pos[1]; // ua < ub, non-wire tile positions
    int e0 = static_cast<int>(threadIdx.x);
    e0 = ((e0 >> ua) << (ua + 1)) | (e0 & ((1 << ua) - 1));
    e0 = ((e0 >> ub) << (ub + 1)) | (e0 & ((1 << ub) - 1));
    double dx = 0.0, dy = 0.0, xr[kTransitionBits], xi[kTransitionBits];
#pragma unroll
    for (int w = 0; w < kTransitionBits; w++)
        xr[w] = xi[w] = 0.0;
    __syncthreads();
    auto issue = [&](uint64_t t, int st) {
        uint64_t base = t; // tile id -> index with zeros at the tile's bit positions
        for (int j = 0; j < tt.n_tile; j++)
            base = insert_zero(base, tt.tile_pos[j]);
        amp_t *sk = stage + size_t(st) * 2 * tile, *sb = sk + tile;
#pragma unroll
        for (int u = 0; u < per_thread; u++) {
            const int e = u * kTransThreads + threadIdx.x;
            const uint64_t gi = base | rowoff[e >> 5] | uint64_t(e & 31);
            cp_async_amp(&sk[e], &ket[gi]);
            cp_async_amp(&sb[e], &bra[gi]);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    int st = 0;
    if (blockIdx.x < n_tiles)
        issue(blockIdx.x, 0);
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, st ^= 1) {
        const uint64_t tn = t + gridDim.x;
        if (tn < n_tiles) {
            issue(tn, st ^ 1);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();
        const amp_t *sk = stage + size_t(st) * 2 * tile, *sb = sk + tile;
#pragma unroll
        for (int u = 0; u < per_thread; u++) {
            const int e = e0 | ((u & 1) << ua) | ((u >> 1) << ub);
            const amp_t h = sb[e], l = sk[e];
            const double cx = double(h.x), cy = -double(h.y); // conj(bra)
            dx = fma(cx, double(l.x), dx);
            dx = fma(-cy, double(l.y), dx);
            dy = fma(cx, double(l.y), dy);
            dy = fma(cy, double(l.x), dy);
#pragma unroll
            for (int w = 0; w < kTransitionBits; w++) {
                if (w < tt.nb) {
                    const amp_t p = sk[e ^ (1 << tt.wire_tpos[w])];
                    xr[w] = fma(cx, double(p.x), xr[w]);
                    xr[w] = fma(-cy, double(p.y), xr[w]);
                    xi[w] = fma(cx, double(p.y), xi[w]);
                    xi[w] = fma(cy, double(p.x), xi[w]);
                }
            }
        }
        __syncthreads(); // the stage is free for the copies issued in the next iteration
    }
    double acc[kTransitionVals];
    acc[0] = dx;
    acc[1] = dy;
#pragma unroll
    for (int w = 0; w < kTransitionBits; w++) {
        const double sg = (w < tt.nb && ((e0 >> tt.wire_tpos[w]) & 1)) ? -1.0 : 1.0;
        const bool on = w < tt.nb;
        acc[2 + 6 * w + 0] = on ? sg * dx : 0.0;
        acc[2 + 6 * w + 1] = on ? sg * dy : 0.0;
        acc[2 + 6 * w + 2] = xr[w];
        acc[2 + 6 * w + 3] = xi[w];
        acc[2 + 6 * w + 4] = sg * xr[w];
        acc[2 + 6 * w + 5] = sg * xi[w];
    }
    block_reduce_store<kTransitionVals>(acc, partials);
}
__global__ void k_finalize_scaled(const double *__restrict__ partials, int nblocks, int nv, int which,
                                  double scale, double *__restrict__ dst) {
    __shared__ double red[kReduceThreads];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        s += partials[b * nv + which];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (static_cast<int>(threadIdx.x) < o)
            red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        *dst = scale * red[0];
}

template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_pauli_sum_apply(const amp_t *__restrict__ in, amp_t *__restrict__ out, uint64_t len,
                      const PauliTerm *__restrict__ terms, int nterms) {
    constexpr int CH = 128;
    __shared__ PauliTerm st[CH];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t i0 = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // all threads of a block iterate the same number of times (len is a multiple of the stride
    // or the tail is guarded), so the __syncthreads below are uniform
    const uint64_t iters = (len + stride - 1) / stride;
    for (uint64_t it = 0; it < iters; it++) {
        const uint64_t i = i0 + it * stride;
        double sr = 0.0, si = 0.0;
        for (int t0 = 0; t0 < nterms; t0 += CH) {
            const int nt = min(CH, nterms - t0);
            __syncthreads();
            if (threadIdx.x < nt)
                st[threadIdx.x] = terms[t0 + threadIdx.x];
            __syncthreads();
            if (i < len) {
                for (int t = 0; t < nt; t++) {
                    const uint64_t j = i ^ st[t].x;
                    const amp_t b = in[j];
                    const double sg = (__popcll(j & st[t].z) & 1) ? -1.0 : 1.0;
                    const double cr = sg * st[t].cr, ci = sg * st[t].ci;
                    sr += cr * b.x - ci * b.y;
                    si += cr * b.y + ci * b.x;
                }
            }
        }
        if (i < len) {
            amp_t o;
            o.x = sr;
            o.y = si;
            out[i] = o;
        }
    }
}

// The same for a sharded state: an X or Y factor on a qubit that sits on a rank bit pairs this shard
// with the shard of rank ^ x_global, so term t gathers from peers.p[st[t].src] -- the partner's shard,
// read in place through its IPC mapping (NVLink loads), no exchange and no staging copy.
template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_pauli_sum_apply_sharded(PeerPtrs peers, amp_t *__restrict__ out, uint64_t len,
                              const PauliTerm *__restrict__ terms, int nterms) {
    constexpr int CH = 128;
    __shared__ PauliTerm st[CH];
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    const uint64_t i0 = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t iters = (len + stride - 1) / stride;
    for (uint64_t it = 0; it < iters; it++) {
        const uint64_t i = i0 + it * stride;
        double sr = 0.0, si = 0.0;
        for (int t0 = 0; t0 < nterms; t0 += CH) {
            const int nt = min(CH, nterms - t0);
            __syncthreads();
            if (threadIdx.x < nt)
                st[threadIdx.x] = terms[t0 + threadIdx.x];
            __syncthreads();
            if (i < len) {
                for (int t = 0; t < nt; t++) {
                    const uint64_t j = i ^ st[t].x;
                    const amp_t b = static_cast<const amp_t *>(peers.p[st[t].src])[j];
                    const double sg = (__popcll(j & st[t].z) & 1) ? -1.0 : 1.0;
                    const double cr = sg * st[t].cr, ci = sg * st[t].ci;
                    sr += cr * b.x - ci * b.y;
                    si += cr * b.y + ci * b.x;
                }
            }
        }
        if (i < len) {
            amp_t o;
            o.x = sr;
            o.y = si;
            out[i] = o;
        }
    }
}

// ---- CSR ----------------------------------------------------------------------------------------
template <typename amp_t, bool EXPVAL>
__global__ void __launch_bounds__(kReduceThreads)
    k_csr(const amp_t *__restrict__ x, amp_t *__restrict__ y, const double2 *__restrict__ data,
          const uint32_t *__restrict__ ind, const uint64_t *__restrict__ ptr, uint64_t nrows,
          int L, double *__restrict__ partials) {
    // L lanes (power of two <= 32) cooperate on one row
    const int sub = threadIdx.x & (L - 1);
    const uint64_t grp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / L;
    const uint64_t ngrp = uint64_t(gridDim.x) * blockDim.x / L;
    double acc[1] = {0.0};
    const uint64_t iters = (nrows + ngrp - 1) / ngrp;
    for (uint64_t it = 0; it < iters; it++) {
        const uint64_t row = grp + it * ngrp;
        double sr = 0.0, si = 0.0;
        if (row < nrows) {
            const uint64_t b = ptr[row], e = ptr[row + 1];
            for (uint64_t j = b + sub; j < e; j += L) {
                const double2 d = data[j];
                const amp_t v = x[ind[j]];
                sr += d.x * v.x - d.y * v.y;
                si += d.x * v.y + d.y * v.x;
            }
        }
        for (int o = L >> 1; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        if (row < nrows && sub == 0) {
            if (EXPVAL) {
                const amp_t a = x[row];
                acc[0] += double(a.x) * sr + double(a.y) * si; // Re(conj(psi_row) * (H psi)_row)
            } else {
                amp_t o;
                o.x = sr;
                o.y = si;
                y[row] = o;
            }
        }
    }
    if (EXPVAL)
        block_reduce_store<1>(acc, partials);
}

// Expectation value of a CSR matrix as a pure stream over the non-zeros:
//   <psi|A|psi> = sum_j Re( conj(psi[row(j)]) * data[j] * psi[ind[j]] ),
// no per-row reduction at all. Every warp owns one contiguous range of the non-zeros: it finds the row
// of its first element by ONE binary search over the row pointers and from then on only advances its
// row cursor (row pointers come from L1); the range is walked in chunks of 32 x K elements whose
// 16-byte and 4-byte loads are all issued before the previous chunk is consumed (register double
// buffering). The gathers psi[ind], psi[row] hit L2 for states up to ~100 MB. Deterministic: every
// lane sums in a fixed order, the block and grid reductions are fixed-order too.
// Sharded states (SHARDED): this rank streams the non-zeros of ITS rows (j_begin .. nnz of the call,
// rows row_begin ..), psi[row] is local and psi[col] comes from whichever shard holds it -- read in
// place through the peers' IPC mappings (csr.n_local = index bits per shard).
struct CsrShard {
    PeerPtrs peers;
    int n_local;
    uint64_t j_begin, row_begin;
};
template <typename amp_t, typename ptr_t, bool SHARDED>
__global__ void __launch_bounds__(kReduceThreads)
    k_csr_expval_stream(const amp_t *__restrict__ x, const double2 *__restrict__ data,
                        const uint32_t *__restrict__ ind, const ptr_t *__restrict__ ptr, uint64_t nrows,
                        uint64_t nnz, CsrShard sh, double *__restrict__ partials) {
    constexpr int K = 4;
    constexpr uint64_t CH = 32 * K;
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    // ranges are whole chunks, so every chunk belongs to exactly one warp
    const uint64_t jb = SHARDED ? sh.j_begin : 0; // the call covers non-zeros jb .. nnz, rows row0 .. nrows
    const uint64_t row0 = SHARDED ? sh.row_begin : 0;
    const uint64_t nchunks = (nnz - jb + CH - 1) / CH;
    const uint64_t per = (nchunks + nwarps - 1) / nwarps;
    const uint64_t c_beg = warp * per, c_end = min(nchunks, c_beg + per);
    const uint64_t local_mask = SHARDED ? ((uint64_t(1) << sh.n_local) - 1) : ~uint64_t(0);
    double acc[1] = {0.0};
    if (c_beg < c_end) {
        uint64_t lo = row0, hi = nrows; // the largest r with ptr[r] <= first element of the range
        const uint64_t jfirst = jb + c_beg * CH;
        while (hi - lo > 1) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (static_cast<uint64_t>(__ldg(ptr + mid)) <= jfirst)
                lo = mid;
            else
                hi = mid;
        }
        uint64_t row = lo;
        uint64_t next_start = static_cast<uint64_t>(__ldg(ptr + row + 1)); // first element of row + 1
        double2 d[K], dn[K];
        uint32_t col[K], coln[K];
        auto load = [&](uint64_t c, double2(&dd)[K], uint32_t(&cc)[K]) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint64_t j = jb + c * CH + lane + 32 * k;
                if (j < nnz) {
                    dd[k] = data[j];
                    cc[k] = ind[j];
                } else {
                    dd[k] = make_double2(0.0, 0.0);
                    cc[k] = 0;
                }
            }
        };
        load(c_beg, d, col);
        for (uint64_t c = c_beg; c < c_end; c++) {
            if (c + 1 < c_end)
                load(c + 1, dn, coln);
            // rows of this lane's K elements first (row pointers come from L1), then all 2K gathers in
            // flight together, then the arithmetic: one L2 round trip per chunk instead of one per element
            uint32_t rowk[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint64_t j = jb + c * CH + lane + 32 * k;
                if (j < nnz) {
                    while (next_start <= j) { // empty rows are skipped the same way
                        row++;
                        next_start = static_cast<uint64_t>(__ldg(ptr + row + 1));
                    }
                }
                rowk[k] = static_cast<uint32_t>(row);
            }
            amp_t v[K], a[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                if constexpr (SHARDED)
                    v[k] = static_cast<const amp_t *>(sh.peers.p[col[k] >> sh.n_local])[col[k] & local_mask];
                else
                    v[k] = x[col[k]];
                a[k] = x[rowk[k] & local_mask];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                // elements past the end carry d = 0
                const double tr = d[k].x * v[k].x - d[k].y * v[k].y, ti = d[k].x * v[k].y + d[k].y * v[k].x;
                acc[0] += double(a[k].x) * tr + double(a[k].y) * ti;
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                d[k] = dn[k];
                col[k] = coln[k];
            }
        }
    }
    block_reduce_store<1>(acc, partials);
}

// ---- probabilities ------------------------------------------------------------------------------
template <typename amp_t>
__global__ void k_probs_full(const amp_t *__restrict__ s, uint64_t len, double *__restrict__ out) {
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t a = s[i];
        out[i] = double(a.x) * a.x + double(a.y) * a.y;
    }
}
struct BitList {
    int m;
    int pos[64];
};
constexpr int kMargSmemBits = 11;
template <typename amp_t, bool PRIVATE>
__global__ void __launch_bounds__(256)
    k_probs_marginal(const amp_t *__restrict__ s, uint64_t len, BitList bl, uint64_t index_or,
                     double *__restrict__ out) {
    __shared__ double hist[PRIVATE ? (1 << kMargSmemBits) : 1];
    const int nb = 1 << bl.m;
    if (PRIVATE) {
        for (int b = threadIdx.x; b < nb; b += blockDim.x)
            hist[b] = 0.0;
        __syncthreads();
    }
    const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const amp_t a = s[i];
        const double p = double(a.x) * a.x + double(a.y) * a.y;
        uint64_t bin = 0;
        const uint64_t gi = i | index_or; // sharded states: the rank supplies the top index bits
        for (int j = 0; j < bl.m; j++)
            bin |= ((gi >> bl.pos[j]) & 1ull) << (bl.m - 1 - j);
        if (PRIVATE)
            atomicAdd(&hist[bin], p);
        else if (p != 0.0)
            atomicAdd(&out[bin], p);
    }
    if (PRIVATE) {
        __syncthreads();
        for (int b = threadIdx.x; b < nb; b += blockDim.x)
            if (hist[b] != 0.0)
                atomicAdd(&out[b], hist[b]);
    }
}

// ---- sampling: chunk sums -> exclusive scan of chunk sums -> per-shot two-level search ----------
template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_chunk_sums(const amp_t *__restrict__ s, uint64_t len, double *__restrict__ chunk) {
    // one block per chunk of 2^kSampleChunkBits amplitudes (or the whole state if smaller)
    const uint64_t csz = uint64_t(1) << kSampleChunkBits;
    const uint64_t b = uint64_t(blockIdx.x) * csz;
    const uint64_t e = (b + csz < len) ? b + csz : len;
    double acc = 0.0;
    for (uint64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const amp_t a = s[i];
        acc += double(a.x) * a.x + double(a.y) * a.y;
    }
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++)
            t += red[w];
        chunk[blockIdx.x] = t;
    }
}
// single block, 1024 threads: in-place exclusive scan of n chunk sums; chunk[n] = total
__global__ void __launch_bounds__(1024) k_scan_chunks(double *__restrict__ chunk, uint64_t n) {
    __shared__ double part[1024];
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t b = uint64_t(threadIdx.x) * per;
    const uint64_t e = (b + per < n) ? b + per : n;
    double sum = 0.0;
    for (uint64_t i = b; i < e; i++)
        sum += chunk[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int t = 0; t < 1024; t++) {
            const double v = part[t];
            part[t] = run;
            run += v;
        }
        chunk[n] = run;
    }
    __syncthreads();
    double run = part[threadIdx.x];
    for (uint64_t i = b; i < e; i++) {
        const double v = chunk[i];
        chunk[i] = run;
        run += v;
    }
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// one warp per shot: inverse-transform sampling, index k with cdf[k] < U <= cdf[k+1]
template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_sample(const amp_t *__restrict__ s, uint64_t len, const double *__restrict__ ccdf,
             uint64_t nchunks, int nq, size_t shots, uint64_t seed, ShardCdf sh,
             unsigned long long *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const size_t shot = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (shot >= shots)
        return;
    // Sharded states: every rank draws the same U in (0, global total]; the rank whose interval
    // (offset, offset + local total] holds it resolves the shot, the others write zeros (the outputs
    // are summed over the ranks afterwards).
    const double total = sh.sharded ? sh.global_total : ccdf[nchunks];
    const uint64_t r = splitmix64(seed ^ splitmix64(shot));
    double U = (double((r >> 11) + 1) * 0x1.0p-53) * total; // (0, total]
    if (sh.sharded) {
        if (!(sh.offset < U && U <= sh.upper)) {
            if (lane == 0)
                for (int j = 0; j < nq; j++)
                    out[shot * nq + j] = 0ull;
            return;
        }
        U -= sh.offset;
    }
    // binary search over chunks: largest c with ccdf[c] < U
    uint64_t lo = 0, hi = nchunks; // invariant: ccdf[lo] < U (ccdf[0] = 0 < U), answer in [lo, hi)
    while (hi - lo > 1) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (ccdf[mid] < U)
            lo = mid;
        else
            hi = mid;
    }
    const uint64_t csz = uint64_t(1) << kSampleChunkBits;
    const uint64_t b = lo * csz;
    const uint64_t e = (b + csz < len) ? b + csz : len;
    double run = ccdf[lo];
    uint64_t found = e - 1;
    bool have = false;
    uint64_t last_nz = b;
    for (uint64_t i0 = b; i0 < e && !have; i0 += 32) {
        const uint64_t i = i0 + lane;
        double p = 0.0;
        if (i < e) {
            const amp_t a = s[i];
            p = double(a.x) * a.x + double(a.y) * a.y;
        }
        double incl = p; // inclusive warp scan
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, (i < e) && (run + incl >= U) && p > 0.0);
        const unsigned nz = __ballot_sync(0xffffffffu, p > 0.0);
        if (hit) {
            found = i0 + (__ffs(hit) - 1);
            have = true;
        } else {
            if (nz)
                last_nz = i0 + (31 - __clz(nz));
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (!have)
        found = last_nz; // rounding at the chunk end: fall back to the last populated entry
    found |= sh.index_or;
    if (lane == 0) {
        for (int j = 0; j < nq; j++) // MSB (wire 0) first, as the reference (MF.hpp:113-115)
            out[shot * nq + (nq - 1 - j)] = (found >> j) & 1ull;
    }
}

// 0/1 sample bits <-> doubles, in place (so that the all-reduce over ranks can sum them)
__global__ void k_u64_to_f64(unsigned long long *p, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        reinterpret_cast<double *>(p)[i] = static_cast<double>(p[i]);
}
__global__ void k_f64_to_u64(unsigned long long *p, size_t n) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = static_cast<unsigned long long>(reinterpret_cast<double *>(p)[i] + 0.5);
}

int reduce_grid(uint64_t work) {
    const uint64_t b = (work + kReduceThreads - 1) / kReduceThreads;
    return static_cast<int>(b < 1 ? 1 : (b > kReduceBlocks ? kReduceBlocks : b));
}
} // namespace

#define DISPATCH_DTYPE(dtype, expr_f, expr_d)                                                   \
    do {                                                                                        \
        if ((dtype) == 1) {                                                                     \
            expr_d;                                                                             \
        } else {                                                                                \
            expr_f;                                                                             \
        }                                                                                       \
        CUDA_CHECK(cudaGetLastError());                                                         \
    } while (0)

// Pairs: local index i with bit l cleared and selector (the top remaining bit) fixed.
//   rank with bit j = 0 holds (0, x): its half with local bit l = 1 trades with the partner's half
//   with local bit l = 0.  my_bit = this rank's bit j; it handles the pairs whose selector bit equals
//   my_bit, so the two kernels touch disjoint pairs and need no synchronisation with each other.
template <typename amp_t>
__global__ void __launch_bounds__(256)
    k_peer_swap(amp_t *__restrict__ mine, amp_t *__restrict__ peer, uint64_t npairs, int lbit,
                int selbit, int my_bit) {
    // Four pairs per thread and iteration: all eight loads (four of them NVLink round trips) are in
    // flight before the first store, which is what keeps the link busy.
    constexpr int U = 4;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const int p0 = lbit < selbit ? lbit : selbit, p1 = lbit < selbit ? selbit : lbit;
    for (uint64_t k0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k0 < npairs;
         k0 += stride * U) {
        uint64_t im[U], ip[U];
        amp_t a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            // k enumerates local indices with bits {lbit, selbit} removed (lbit != selbit)
            const uint64_t k = k0 + u * stride;
            uint64_t i = insert_zero(insert_zero(k < npairs ? k : k0, p0), p1);
            i |= static_cast<uint64_t>(my_bit) << selbit;
            im[u] = i | (static_cast<uint64_t>(1 - my_bit) << lbit); // my half to give away
            ip[u] = i | (static_cast<uint64_t>(my_bit) << lbit);     // partner's half
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            b[u] = peer[ip[u]];
            a[u] = mine[im[u]];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + u * stride < npairs) {
                peer[ip[u]] = a[u];
                mine[im[u]] = b[u];
            }
        }
    }
}
void launch_peer_swap(int dtype, void *mine, void *peer, int n_local, int lbit, int my_bit,
                      cudaStream_t st) {
    // selector: the highest local bit that is not lbit (keeps each warp on one contiguous run)
    const int selbit = (lbit == n_local - 1) ? n_local - 2 : n_local - 1;
    const uint64_t npairs = uint64_t(1) << (n_local - 2);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((npairs + 255) / 256, 148 * 16));
    if (dtype == 1)
        k_peer_swap<double2><<<grid, 256, 0, st>>>(static_cast<double2 *>(mine),
                                                   static_cast<double2 *>(peer), npairs, lbit,
                                                   selbit, my_bit);
    else
        k_peer_swap<float2><<<grid, 256, 0, st>>>(static_cast<float2 *>(mine),
                                                  static_cast<float2 *>(peer), npairs, lbit, selbit,
                                                  my_bit);
    CUDA_CHECK(cudaGetLastError());
}

// ---- k-bit global<->local exchange: an all-to-all inside a group of 2^k ranks, in place ---------
// Rank bits j_1..j_k trade places with local bits l_1..l_k. With a = this rank's value on the rank
// bits and b != a a partner's, the sub-block of this shard whose local bits spell b and the sub-block
// of the partner's shard whose local bits spell a swap contents element by element (the sub-block
// with local bits = a stays). For every unordered pair {a, b} each of the two ranks moves half of the
// elements (selector bit = a < b on one side, b < a on the other), reading both sides and writing
// both sides through the partner's IPC-mapped shard, so every NVLink direction of every rank carries
// (1 - 2^-k) S / 2 of loads and as much of stores. Chunks of 1024 elements go round-robin over the
// partners, starting at a + 1, so at any moment the ranks of a group talk to distinct peers.
template <typename amp_t, int T>
__global__ void __launch_bounds__(T)
    k_exchange(amp_t *__restrict__ mine, ExchangeParams p, uint64_t nrest, uint64_t nchunks) {
    constexpr int U = 4;
    const uint32_t nparts = (1u << p.k) - 1u;
    for (uint64_t c = blockIdx.x; c < nchunks * nparts; c += gridDim.x) {
        const uint32_t pi = static_cast<uint32_t>(c % nparts);
        const uint64_t chunk = c / nparts;
        const uint32_t b = (p.a + 1u + pi) & nparts; // nparts = 2^k - 1 is also the group mask
        amp_t *__restrict__ peer = static_cast<amp_t *>(p.peer[b]);
        uint64_t dep_a = 0, dep_b = 0;
#pragma unroll 4
        for (int i = 0; i < p.k; i++) {
            dep_a |= static_cast<uint64_t>((p.a >> i) & 1u) << p.lpos[i];
            dep_b |= static_cast<uint64_t>((b >> i) & 1u) << p.lpos[i];
        }
        const uint64_t sel = static_cast<uint64_t>(p.a < b ? 0 : 1) << p.selbit;
        uint64_t im[U], ip[U];
        amp_t x[U], y[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = chunk * (T * U) + u * T + threadIdx.x;
            ok[u] = e < nrest;
            uint64_t i = ok[u] ? e : 0;
            for (int f = 0; f < p.nfix; f++)
                i = insert_zero(i, p.fixpos[f]);
            i |= sel;
            im[u] = i | dep_b;
            ip[u] = i | dep_a;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            y[u] = peer[ip[u]];
            x[u] = mine[im[u]];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (ok[u]) {
                peer[ip[u]] = x[u];
                mine[im[u]] = y[u];
            }
        }
    }
}
void launch_exchange(int dtype, void *mine, const ExchangeParams &p, int max_ctas, bool fat,
                     cudaStream_t st) {
    const uint64_t nrest = uint64_t(1) << p.nfree;
    const uint64_t per_chunk = (fat ? 1024 : 256) * 4;
    const uint64_t nchunks = (nrest + per_chunk - 1) / per_chunk;
    const uint64_t work = nchunks * ((uint64_t(1) << p.k) - 1);
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(work, static_cast<uint64_t>(max_ctas)));
    if (dtype == 1) {
        if (fat)
            k_exchange<double2, 1024><<<grid, 1024, 0, st>>>(static_cast<double2 *>(mine), p, nrest, nchunks);
        else
            k_exchange<double2, 256><<<grid, 256, 0, st>>>(static_cast<double2 *>(mine), p, nrest, nchunks);
    } else {
        if (fat)
            k_exchange<float2, 1024><<<grid, 1024, 0, st>>>(static_cast<float2 *>(mine), p, nrest, nchunks);
        else
            k_exchange<float2, 256><<<grid, 256, 0, st>>>(static_cast<float2 *>(mine), p, nrest, nchunks);
    }
    CUDA_CHECK(cudaGetLastError());
}

// ---- cross-rank barrier through IPC-mapped flag words (one block, one thread per rank) ------------
// Thread r publishes `epoch` into rank r's flag array (slot = this rank) and then waits until rank r's
// value in this rank's own array reaches `epoch`. Stream order makes everything queued before the
// barrier on this rank complete (remote stores included) before the flag goes out. A peer that never
// arrives must not hang the GPU: after `timeout_ns` the kernel traps (a CUDA error on the host).
__global__ void k_flag_barrier(unsigned long long *mine, FlagPeers peers, int rank, int world,
                               unsigned long long epoch, unsigned long long timeout_ns) {
    const int r = threadIdx.x;
    if (r >= world)
        return;
    __threadfence_system();
    unsigned long long *dst = peers.p[r] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + r) : "memory");
        if (v >= epoch)
            break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns)
            __trap();
    }
    __threadfence_system();
}
void launch_flag_barrier(unsigned long long *mine, const FlagPeers &peers, int rank, int world,
                         unsigned long long epoch, cudaStream_t st) {
    k_flag_barrier<<<1, 32, 0, st>>>(mine, peers, rank, world, epoch, 20ull * 1000 * 1000 * 1000);
    CUDA_CHECK(cudaGetLastError());
}

void launch_set_basis(int dtype, void *state, uint64_t len, uint64_t index, cudaStream_t st) {
    const int grid = reduce_grid(len);
    DISPATCH_DTYPE(dtype,
                   (k_set_basis<float2><<<grid, 256, 0, st>>>(static_cast<float2 *>(state), len, index)),
                   (k_set_basis<double2><<<grid, 256, 0, st>>>(static_cast<double2 *>(state), len, index)));
}
void launch_scatter(int dtype, void *state, const uint64_t *d_idx, const double2 *d_val, size_t n,
                    cudaStream_t st) {
    if (n == 0)
        return;
    const int grid = static_cast<int>((n + 255) / 256);
    DISPATCH_DTYPE(dtype,
                   (k_scatter<float2><<<grid, 256, 0, st>>>(static_cast<float2 *>(state), d_idx, d_val, n)),
                   (k_scatter<double2><<<grid, 256, 0, st>>>(static_cast<double2 *>(state), d_idx, d_val, n)));
}
void launch_scatter_wires(int dtype, void *state, const double2 *d_val, const int *h_pos, int k,
                          int n_local, uint64_t rank, cudaStream_t st) {
    WirePos wp{};
    wp.k = k;
    for (int j = 0; j < k; j++)
        wp.pos[j] = h_pos[j];
    const uint64_t count = uint64_t(1) << k;
    const int grid = reduce_grid(count);
    DISPATCH_DTYPE(dtype,
                   (k_scatter_wires<float2><<<grid, 256, 0, st>>>(static_cast<float2 *>(state), d_val, count, wp, n_local, rank)),
                   (k_scatter_wires<double2><<<grid, 256, 0, st>>>(static_cast<double2 *>(state), d_val, count, wp, n_local, rank)));
}
void launch_gather(int dtype, const void *state, const uint64_t *d_idx, size_t n, int n_local,
                   uint64_t rank, double2 *d_out, cudaStream_t st) {
    if (n == 0)
        return;
    const int grid = static_cast<int>((n + 255) / 256);
    DISPATCH_DTYPE(dtype,
                   (k_gather<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), d_idx, n, n_local, rank, d_out)),
                   (k_gather<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), d_idx, n, n_local, rank, d_out)));
}
void launch_axpy(int dtype, double ar, double ai, const void *x, void *y, uint64_t len,
                 cudaStream_t st) {
    const int grid = reduce_grid(len);
    DISPATCH_DTYPE(dtype,
                   (k_axpy<float2, float><<<grid, 256, 0, st>>>(float(ar), float(ai), static_cast<const float2 *>(x), static_cast<float2 *>(y), len)),
                   (k_axpy<double2, double><<<grid, 256, 0, st>>>(ar, ai, static_cast<const double2 *>(x), static_cast<double2 *>(y), len)));
}
void launch_matk(int dtype, void *state, int n_eff, const double2 *d_mat, const int *h_bits, int k,
                 cudaStream_t st) {
    MatKParams p{};
    p.k = k;
    for (int j = 0; j < k; j++)
        p.bits[j] = p.sorted[j] = h_bits[j];
    for (int a = 0; a < k; a++)
        for (int b = a + 1; b < k; b++)
            if (p.sorted[b] < p.sorted[a]) {
                const int t = p.sorted[a];
                p.sorted[a] = p.sorted[b];
                p.sorted[b] = t;
            }
    const int dim = 1 << k;
    const int threads = 128;
    const int G = threads >= dim ? threads / dim : 1;
    const uint64_t ngroups = uint64_t(1) << (n_eff - k);
    const uint64_t niter = (ngroups + G - 1) / G;
    const int grid = static_cast<int>(niter < 148 * 16 ? niter : 148 * 16);
    const size_t smem = sizeof(double2) * size_t(G) * dim;
    DISPATCH_DTYPE(dtype,
                   (k_matk<float2><<<grid, threads, smem, st>>>(static_cast<float2 *>(state), d_mat, p, ngroups)),
                   (k_matk<double2><<<grid, threads, smem, st>>>(static_cast<double2 *>(state), d_mat, p, ngroups)));
}

void launch_norm2(int dtype, const void *state, uint64_t len, double *d_partials, cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_norm2<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), len, d_partials)),
                   (k_norm2<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), len, d_partials)));
}
void launch_dot(int dtype, const void *x, const void *y, uint64_t len, double *d_partials,
                cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_dot<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(x), static_cast<const float2 *>(y), len, d_partials)),
                   (k_dot<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(x), static_cast<const double2 *>(y), len, d_partials)));
}
void launch_expval_1q(int dtype, const void *state, uint64_t len, int tbit, const double *m8,
                      double *d_partials, cudaStream_t st) {
    M8 mm;
    for (int i = 0; i < 8; i++)
        mm.m[i] = m8[i];
    DISPATCH_DTYPE(dtype,
                   (k_expval_1q<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), len / 2, tbit, mm, d_partials)),
                   (k_expval_1q<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), len / 2, tbit, mm, d_partials)));
}
void launch_expval_2q(int dtype, const void *state, uint64_t len, int bit_a, int bit_b,
                      const double2 *d_m16, double *d_partials, cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_expval_2q<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), len / 4, bit_a, bit_b, d_m16, d_partials)),
                   (k_expval_2q<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), len / 4, bit_a, bit_b, d_m16, d_partials)));
}
void launch_pauli_expval(int dtype, const void *state, uint64_t len, uint64_t x, uint64_t z,
                         double phr, double phi, double *d_partials, cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_pauli_expval<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), len, x, z, phr, phi, d_partials)),
                   (k_pauli_expval<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), len, x, z, phr, phi, d_partials)));
}
void launch_pauli_dot(int dtype, const void *bra, const void *ket, uint64_t len, uint64_t x,
                      uint64_t z, double phr, double phi, double *d_partials, cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_pauli_dot<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(bra), static_cast<const float2 *>(ket), len, x, z, phr, phi, d_partials)),
                   (k_pauli_dot<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(bra), static_cast<const double2 *>(ket), len, x, z, phr, phi, d_partials)));
}
void launch_expval_z_all(int dtype, const void *state, int n_bits, double *d_partials, cudaStream_t st) {
    B2_ASSERT(n_bits >= 12 && n_bits <= kZAllMaxBits);
    const uint64_t nchunks = uint64_t(1) << (n_bits - 12);
    const int grid = static_cast<int>(std::min<uint64_t>(nchunks, kReduceBlocks));
    if (grid < kReduceBlocks) // finalize sums kReduceBlocks rows
        CUDA_CHECK(cudaMemsetAsync(d_partials, 0, sizeof(double) * kReduceBlocks * kZAllValsHost, st));
    DISPATCH_DTYPE(dtype,
                   (k_expval_z_all<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), n_bits, d_partials)),
                   (k_expval_z_all<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), n_bits, d_partials)));
}
void launch_pauli_sum_expval(int dtype, const void *state, uint64_t len, const PauliTerm *d_terms,
                             int nterms, double *d_partials, cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_pauli_sum_expval<float2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), len, d_terms, nterms, d_partials)),
                   (k_pauli_sum_expval<double2><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), len, d_terms, nterms, d_partials)));
}
void launch_finalize_scaled(const double *d_partials, int nblocks, int nv, int which, double scale,
                            double *d_dst, cudaStream_t st) {
    k_finalize_scaled<<<1, kReduceThreads, 0, st>>>(d_partials, nblocks, nv, which, scale, d_dst);
    CUDA_CHECK(cudaGetLastError());
}
void launch_finalize(const double *d_partials, int nblocks, int nv, double *d_out,
                     cudaStream_t st) {
    if (nv > 4)
        k_finalize_wide<<<nv, kReduceThreads, 0, st>>>(d_partials, nblocks, nv, d_out);
    else
        k_finalize<<<1, kReduceThreads, 0, st>>>(d_partials, nblocks, nv, d_out);
    CUDA_CHECK(cudaGetLastError());
}
void launch_pauli_sum_apply(int dtype, const void *in, void *out, uint64_t len,
                            const PauliTerm *d_terms, int nterms, cudaStream_t st) {
    const uint64_t b = (len + 255) / 256;
    const int grid = static_cast<int>(b < 148 * 8 ? b : 148 * 8);
    DISPATCH_DTYPE(dtype,
                   (k_pauli_sum_apply<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(in), static_cast<float2 *>(out), len, d_terms, nterms)),
                   (k_pauli_sum_apply<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(in), static_cast<double2 *>(out), len, d_terms, nterms)));
}
void launch_pauli_sum_apply_sharded(int dtype, const PeerPtrs &peers, void *out, uint64_t len,
                                    const PauliTerm *d_terms, int nterms, cudaStream_t st) {
    const uint64_t b = (len + 255) / 256;
    const int grid = static_cast<int>(b < 148 * 8 ? b : 148 * 8);
    DISPATCH_DTYPE(dtype,
                   (k_pauli_sum_apply_sharded<float2><<<grid, 256, 0, st>>>(peers, static_cast<float2 *>(out), len, d_terms, nterms)),
                   (k_pauli_sum_apply_sharded<double2><<<grid, 256, 0, st>>>(peers, static_cast<double2 *>(out), len, d_terms, nterms)));
}
void launch_csr_expval_stream(int dtype, const void *state, const double2 *d_data, const uint32_t *d_ind,
                              const uint64_t *d_ptr64, const uint32_t *d_ptr32, uint64_t nrows, uint64_t nnz,
                              double *d_partials, cudaStream_t st) {
    const CsrShard none{};
    if (d_ptr32) {
        DISPATCH_DTYPE(dtype,
                       (k_csr_expval_stream<float2, uint32_t, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), d_data, d_ind, d_ptr32, nrows, nnz, none, d_partials)),
                       (k_csr_expval_stream<double2, uint32_t, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), d_data, d_ind, d_ptr32, nrows, nnz, none, d_partials)));
    } else {
        DISPATCH_DTYPE(dtype,
                       (k_csr_expval_stream<float2, uint64_t, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), d_data, d_ind, d_ptr64, nrows, nnz, none, d_partials)),
                       (k_csr_expval_stream<double2, uint64_t, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), d_data, d_ind, d_ptr64, nrows, nnz, none, d_partials)));
    }
}
// rows row_begin .. row_end of the matrix belong to this shard; non-zeros j_begin .. j_end
void launch_csr_expval_sharded(int dtype, const void *state, const PeerPtrs &peers, int n_local,
                               const double2 *d_data, const uint32_t *d_ind, const uint64_t *d_ptr64,
                               uint64_t row_begin, uint64_t row_end, uint64_t j_begin, uint64_t j_end,
                               double *d_partials, cudaStream_t st) {
    CsrShard sh{};
    sh.peers = peers;
    sh.n_local = n_local;
    sh.j_begin = j_begin;
    sh.row_begin = row_begin;
    DISPATCH_DTYPE(dtype,
                   (k_csr_expval_stream<float2, uint64_t, true><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), d_data, d_ind, d_ptr64, row_end, j_end, sh, d_partials)),
                   (k_csr_expval_stream<double2, uint64_t, true><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), d_data, d_ind, d_ptr64, row_end, j_end, sh, d_partials)));
}
void launch_csr_expval(int dtype, const void *state, const double2 *d_data, const uint32_t *d_ind,
                       const uint64_t *d_ptr, uint64_t nrows, int L, double *d_partials,
                       cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_csr<float2, true><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(state), nullptr, d_data, d_ind, d_ptr, nrows, L, d_partials)),
                   (k_csr<double2, true><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(state), nullptr, d_data, d_ind, d_ptr, nrows, L, d_partials)));
}
void launch_csr_spmv(int dtype, const void *x, void *y, const double2 *d_data,
                     const uint32_t *d_ind, const uint64_t *d_ptr, uint64_t nrows, int L,
                     cudaStream_t st) {
    DISPATCH_DTYPE(dtype,
                   (k_csr<float2, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(x), static_cast<float2 *>(y), d_data, d_ind, d_ptr, nrows, L, nullptr)),
                   (k_csr<double2, false><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(x), static_cast<double2 *>(y), d_data, d_ind, d_ptr, nrows, L, nullptr)));
}
void launch_probs_full(int dtype, const void *state, uint64_t len, double *d_out, cudaStream_t st) {
    const int grid = reduce_grid(len);
    DISPATCH_DTYPE(dtype,
                   (k_probs_full<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), len, d_out)),
                   (k_probs_full<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), len, d_out)));
}
void launch_probs_marginal(int dtype, const void *state, uint64_t len, const int *h_bitpos, int m,
                           uint64_t index_or, double *d_out, cudaStream_t st) {
    BitList bl{};
    bl.m = m;
    for (int j = 0; j < m; j++)
        bl.pos[j] = h_bitpos[j];
    const int grid = reduce_grid(len);
    if (m <= kMargSmemBits) {
        DISPATCH_DTYPE(dtype,
                       (k_probs_marginal<float2, true><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), len, bl, index_or, d_out)),
                       (k_probs_marginal<double2, true><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), len, bl, index_or, d_out)));
    } else {
        DISPATCH_DTYPE(dtype,
                       (k_probs_marginal<float2, false><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), len, bl, index_or, d_out)),
                       (k_probs_marginal<double2, false><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), len, bl, index_or, d_out)));
    }
}
void launch_chunk_sums(int dtype, const void *state, uint64_t len, double *d_chunk,
                       cudaStream_t st) {
    const uint64_t csz = uint64_t(1) << kSampleChunkBits;
    const int grid = static_cast<int>((len + csz - 1) / csz);
    DISPATCH_DTYPE(dtype,
                   (k_chunk_sums<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), len, d_chunk)),
                   (k_chunk_sums<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), len, d_chunk)));
}
void launch_scan_chunks(double *d_chunk, uint64_t nchunks, cudaStream_t st) {
    k_scan_chunks<<<1, 1024, 0, st>>>(d_chunk, nchunks);
    CUDA_CHECK(cudaGetLastError());
}
void launch_sample(int dtype, const void *state, uint64_t len, const double *d_chunk_cdf,
                   uint64_t nchunks, int num_qubits, size_t shots, uint64_t seed,
                   const ShardCdf &sh, unsigned long long *d_out, cudaStream_t st) {
    if (shots == 0)
        return;
    const int grid = static_cast<int>((shots * 32 + 255) / 256);
    DISPATCH_DTYPE(dtype,
                   (k_sample<float2><<<grid, 256, 0, st>>>(static_cast<const float2 *>(state), len, d_chunk_cdf, nchunks, num_qubits, shots, seed, sh, d_out)),
                   (k_sample<double2><<<grid, 256, 0, st>>>(static_cast<const double2 *>(state), len, d_chunk_cdf, nchunks, num_qubits, shots, seed, sh, d_out)));
}

void launch_transition_1q(int dtype, const void *bra, const void *ket, uint64_t len,
                          const int *h_bits, int nb, double *d_partials, cudaStream_t st) {
    B2_ASSERT(nb >= 1 && nb <= kTransitionBits);
    TransBits tb{};
    tb.nb = nb;
    for (int j = 0; j < nb; j++)
        tb.pos[j] = h_bits[j];
    DISPATCH_DTYPE(dtype,
                   (k_transition_1q<float2, kTransitionBits><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const float2 *>(bra), static_cast<const float2 *>(ket), len, tb, d_partials)),
                   (k_transition_1q<double2, kTransitionBits><<<kReduceBlocks, kReduceThreads, 0, st>>>(static_cast<const double2 *>(bra), static_cast<const double2 *>(ket), len, tb, d_partials)));
}
void launch_transition_tile(int dtype, const void *bra, const void *ket, int n_bits,
                            const int *h_bits, int nb, double *d_partials, cudaStream_t st) {
    B2_ASSERT(nb >= 1 && nb <= kTransitionBits && n_bits >= 11);
    // tile bits: 0..4 plus the requested bits >= 5, padded with the lowest unused bits up to 11
    uint64_t mask = 0x1f;
    for (int j = 0; j < nb; j++)
        mask |= uint64_t(1) << h_bits[j];
    for (int b = 5; __builtin_popcountll(mask) < 11; b++)
        mask |= uint64_t(1) << b;
    TransTile tt{};
    tt.nb = nb;
    for (int b = 0; b < 64; b++)
        if ((mask >> b) & 1)
            tt.tile_pos[tt.n_tile++] = b;
    B2_ASSERT(tt.n_tile == 11 && tt.tile_pos[10] < n_bits);
    uint32_t wire_positions = 0;
    for (int j = 0; j < nb; j++)
        for (int k = 0; k < tt.n_tile; k++)
            if (tt.tile_pos[k] == h_bits[j]) {
                tt.wire_tpos[j] = k;
                wire_positions |= 1u << k;
            }
    { // the two highest non-wire tile positions (at most 6 of the 11 positions are wires)
        int found = 0;
        for (int k = tt.n_tile - 1; k >= 0 && found < 2; k--)
            if (!((wire_positions >> k) & 1u))
                tt.u_pos[1 - found++] = k;
        B2_ASSERT(found == 2);
    }
    const uint64_t n_tiles = uint64_t(1) << (n_bits - tt.n_tile);
    const int sms = sm_count_current_device();
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n_tiles, static_cast<uint64_t>(sms)));
    const bool f32 = dtype != 1;
    const size_t smem = (f32 ? sizeof(float2) : sizeof(double2)) * 4 * (size_t(1) << tt.n_tile); // 2 stages x 2 vectors
    static uint64_t configured = 0; // one bit per device
    if (first_use_on_device(configured)) {
        CUDA_CHECK(cudaFuncSetAttribute(k_transition_tile<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        CUDA_CHECK(cudaFuncSetAttribute(k_transition_tile<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    }
    if (grid < static_cast<unsigned>(kReduceBlocks)) // finalize sums kReduceBlocks rows
        CUDA_CHECK(cudaMemsetAsync(d_partials, 0, sizeof(double) * kReduceBlocks * kTransitionVals, st));
    DISPATCH_DTYPE(dtype,
                   (k_transition_tile<float2><<<grid, kTransThreads, smem, st>>>(static_cast<const float2 *>(bra), static_cast<const float2 *>(ket), n_tiles, tt, d_partials)),
                   (k_transition_tile<double2><<<grid, kTransThreads, smem, st>>>(static_cast<const double2 *>(bra), static_cast<const double2 *>(ket), n_tiles, tt, d_partials)));
}
void launch_bits_to_f64(unsigned long long *d, size_t n, cudaStream_t st) {
    if (n)
        k_u64_to_f64<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d, n);
    CUDA_CHECK(cudaGetLastError());
}
void launch_f64_to_bits(unsigned long long *d, size_t n, cudaStream_t st) {
    if (n)
        k_f64_to_u64<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d, n);
    CUDA_CHECK(cudaGetLastError());
}

} // namespace b2sv
