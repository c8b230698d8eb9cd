// b2sv: the extern "C" boundary (include/b2sv.h). Thin: argument marshalling + error capture.
#include "../../include/b2sv.h"

#include "adjoint.hpp"
#include "comm.hpp"
#include "obs.hpp"
#include "shard_plan.hpp"
#include "state.hpp"

#include <iomanip>

#include <cstring>

using namespace b2sv;

struct b2sv_state {
    std::unique_ptr<State> s;
};
struct b2sv_obs {
    ObsPtr o;
};
struct b2sv_ops {
    OpsData d;
    mutable std::shared_ptr<PlanCache> plan; // schedule (+ CUDA graph) of the last state geometry it ran on
};
struct b2sv_csr {
    std::shared_ptr<CsrDevice> m;
};

namespace {
thread_local std::string g_err;

template <class F> int guard(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}
std::vector<int64_t> wires_vec(const int64_t *w, int nw) {
    B2_ABORT_IF(nw < 0 || (nw > 0 && !w), "invalid wires argument");
    return std::vector<int64_t>(w, w + nw);
}
std::vector<cplx> cplx_vec(const double *p, size_t n) {
    B2_ABORT_IF(n > 0 && !p, "invalid complex buffer argument");
    std::vector<cplx> v(n);
    for (size_t i = 0; i < n; i++)
        v[i] = cplx(p[2 * i], p[2 * i + 1]);
    return v;
}
State &st(b2sv_state *s) {
    B2_ABORT_IF(!s || !s->s, "null state handle");
    return *s->s;
}
const State &st(const b2sv_state *s) {
    B2_ABORT_IF(!s || !s->s, "null state handle");
    return *s->s;
}
// number of complex entries of a matrix on nw wires; nw is bounded by what lower_matrix accepts, so a
// bogus count can neither shift out of range nor make us read far beyond the caller's buffer
constexpr int kMaxMatrixWires = 10;
size_t matrix_len(int nw) {
    B2_ABORT_IF(nw < 1 || nw > kMaxMatrixWires,
                "matrix operations support between 1 and 10 wires");
    return size_t(1) << (2 * nw);
}
std::string name_str(const char *name) {
    B2_ABORT_IF(!name, "null operation name");
    return std::string(name);
}
void copy_str(const std::string &s, char *buf, size_t cap) {
    B2_ABORT_IF(!buf || cap == 0, "invalid string buffer");
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
}
} // namespace

extern "C" {

const char *b2sv_last_error(void) { return g_err.c_str(); }
const char *b2sv_version(void) { return "b2sv 0.1.0 (sm_100a)"; }

int b2sv_device_count(int *count) {
    return guard([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            n = 0;
            cudaGetLastError();
        }
        *count = n;
    });
}
int b2sv_backend_info(char *buf, size_t cap) {
    return guard([&] {
        std::ostringstream os;
        os << "backend: b2sv hand-written CUDA (sm_100a), no Kokkos\n";
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            n = 0;
            cudaGetLastError();
        }
        os << "devices: " << n << "\n";
        for (int d = 0; d < n; d++) {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, d) == cudaSuccess)
                os << "  [" << d << "] " << p.name << " sm_" << p.major << p.minor << " SMs="
                   << p.multiProcessorCount << " mem=" << (p.totalGlobalMem >> 20) << " MiB\n";
        }
        int B, R;
        tile_config(B2SV_C128, &B, &R);
        os << "tile: c128 2^" << B << " amplitudes, 2^" << R << " per thread\n";
        copy_str(os.str(), buf, cap);
    });
}

int b2sv_create(int num_qubits, int dtype, int device_id, b2sv_state **out) {
    return guard([&] {
        B2_ABORT_IF(!out, "null output pointer");
        auto h = std::make_unique<b2sv_state>();
        h->s = std::make_unique<State>(num_qubits, dtype, device_id);
        *out = h.release();
    });
}
int b2sv_create_sharded(int num_qubits_total, int dtype, int device_id, int rank, int world,
                        const void *nccl_unique_id, b2sv_state **out) {
    return guard([&] {
        B2_ABORT_IF(!out, "null output pointer");
        B2_ABORT_IF(world > 1 && !nccl_unique_id, "sharded state needs an NCCL unique id");
        auto h = std::make_unique<b2sv_state>();
        h->s = std::make_unique<State>(num_qubits_total, dtype, device_id, rank, world, nccl_unique_id);
        *out = h.release();
    });
}
int b2sv_comm_unique_id(void *out128) {
    return guard([&] { comm_unique_id(out128); });
}
int b2sv_destroy(b2sv_state *s) {
    return guard([&] { delete s; });
}
int b2sv_clone(const b2sv_state *src, b2sv_state **out) {
    return guard([&] {
        auto h = std::make_unique<b2sv_state>();
        h->s = st(src).clone();
        *out = h.release();
    });
}
int b2sv_copy(b2sv_state *dst, const b2sv_state *src) {
    return guard([&] { st(dst).copy_from(st(src)); });
}
int b2sv_reset(b2sv_state *s) {
    return guard([&] { st(s).reset(); });
}
int b2sv_init_zeros(b2sv_state *s) {
    return guard([&] { st(s).init_zeros(); });
}
int b2sv_set_basis_state(b2sv_state *s, uint64_t index) {
    return guard([&] { st(s).set_basis_state(index); });
}
int b2sv_set_state_vector(b2sv_state *s, const uint64_t *indices, const double *values, size_t n) {
    return guard([&] {
        auto v = cplx_vec(values, n);
        st(s).set_state_vector(indices, v.data(), n);
    });
}
int b2sv_set_state_on_wires(b2sv_state *s, const int64_t *wires, int nw, const double *values) {
    return guard([&] {
        B2_ABORT_IF(nw < 1 || nw > 40 || !values, "invalid arguments");
        auto v = cplx_vec(values, size_t(1) << nw);
        st(s).set_state_on_wires(wires_vec(wires, nw), v.data());
    });
}
int b2sv_h2d(b2sv_state *s, const void *host, size_t length) {
    return guard([&] { st(s).h2d(host, length); });
}
int b2sv_d2h(const b2sv_state *s, void *host, size_t length) {
    return guard([&] { st(s).d2h(host, length); });
}
int b2sv_get_amplitudes(const b2sv_state *s, const uint64_t *indices, size_t n, double *out) {
    return guard([&] {
        B2_ABORT_IF(n > 0 && (!indices || !out), "null buffer");
        st(s).get_amplitudes(indices, n, reinterpret_cast<cplx *>(out));
    });
}
int b2sv_trace_begin(b2sv_state *s) {
    return guard([&] { st(s).trace_begin(); });
}
int b2sv_trace_end(b2sv_state *s, int *kinds, double *start_ms, double *dur_ms, int cap, int *n) {
    return guard([&] {
        B2_ABORT_IF(!n, "null output");
        const auto recs = st(s).trace_end();
        *n = static_cast<int>(recs.size());
        for (int i = 0; i < cap && i < *n; i++) {
            if (kinds)
                kinds[i] = recs[i].kind;
            if (start_ms)
                start_ms[i] = recs[i].start_ms;
            if (dur_ms)
                dur_ms[i] = recs[i].dur_ms;
        }
    });
}
int b2sv_num_qubits(const b2sv_state *s, int *n) {
    return guard([&] { *n = st(s).num_qubits(); });
}
int b2sv_data_length(const b2sv_state *s, uint64_t *len) {
    return guard([&] { *len = st(s).local_length(); });
}
int b2sv_device_ptr(const b2sv_state *s, void **ptr) {
    return guard([&] {
        *ptr = st(s).data();
        // the caller may write through the pointer: cached measurements are dropped now (and must be
        // dropped again with b2sv_invalidate after every later write through a retained pointer)
        const_cast<State &>(st(s)).touch();
    });
}
int b2sv_invalidate(b2sv_state *s) {
    return guard([&] { st(s).touch(); });
}
int b2sv_expval_z_all(const b2sv_state *s, double *out, int cap) {
    return guard([&] {
        B2_ABORT_IF(!out, "null output");
        const State &sv = st(s);
        if (sv.num_local() >= 12 && sv.num_local() <= 40 && sv.alloc_length() == sv.local_length()) {
            const std::vector<double> &z = sv.expval_z_all();
            for (int w = 0; w < cap && w < sv.num_qubits(); w++)
                out[w] = z[w];
        } else {
            for (int w = 0; w < cap && w < sv.num_qubits(); w++)
                out[w] = sv.expval_named("PauliZ", {static_cast<int64_t>(w)});
        }
    });
}
int b2sv_stream(const b2sv_state *s, void **stream) {
    return guard([&] { *stream = static_cast<void *>(st(s).stream()); });
}
int b2sv_sync(const b2sv_state *s) {
    return guard([&] { st(s).sync(); });
}

int b2sv_apply(b2sv_state *s, const char *name, const int64_t *wires, int nw, int inverse,
               const double *params, int np) {
    return guard([&] {
        GateOp op;
        op.name = name_str(name);
        op.wires = wires_vec(wires, nw);
        op.inverse = inverse != 0;
        B2_ABORT_IF(np < 0 || (np > 0 && !params), "invalid params argument");
        if (np > 0)
            op.params.assign(params, params + np);
        B2_ABORT_IF(op.name != "Identity" && !is_named_gate(op.name),
                    "operation '" + op.name + "' is not a named gate; use b2sv_apply_matrix");
        st(s).apply_gate(op);
    });
}
int b2sv_apply_matrix(b2sv_state *s, const int64_t *wires, int nw, int inverse,
                      const double *matrix) {
    return guard([&] {
        GateOp op;
        op.name = "__matrix__";
        op.wires = wires_vec(wires, nw);
        op.inverse = inverse != 0;
        op.matrix = cplx_vec(matrix, matrix_len(nw));
        st(s).apply_gate(op);
    });
}
int b2sv_apply_ops(b2sv_state *s, const b2sv_ops *ops, int adjoint) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null ops handle");
        if (adjoint)
            st(s).apply_ops(ops->d.ops, true);
        else
            st(s).apply_ops_cached(ops->d.ops, ops->plan);
    });
}
int b2sv_apply_generator(b2sv_state *s, const char *name, const int64_t *wires, int nw, int adj,
                         double *scale) {
    (void)adj; // ignored by every generator of the reference
    return guard([&] { *scale = st(s).apply_generator(name_str(name), wires_vec(wires, nw)); });
}
int b2sv_set_fusion(b2sv_state *s, int fuse) {
    return guard([&] { st(s).set_fusion(fuse != 0); });
}
int b2sv_get_stats(const b2sv_state *s, uint64_t *sweeps, uint64_t *launches) {
    return guard([&] {
        if (sweeps)
            *sweeps = st(s).sweeps;
        if (launches)
            *launches = st(s).launches + st(s).reduce_launches;
    });
}
int b2sv_last_adjoint_traffic(const b2sv_state *s, uint64_t *bytes) {
    return guard([&] {
        B2_ABORT_IF(!bytes, "null output");
        *bytes = st(s).last_adjoint_bytes;
    });
}
int b2sv_reset_stats(b2sv_state *s) {
    return guard([&] {
        st(s).sweeps = 0;
        st(s).launches = 0;
        st(s).reduce_launches = 0;
    });
}
int b2sv_debug_tile_prof(uint64_t *out16) {
    return guard([&] {
        B2_ABORT_IF(!out16, "null output");
        unsigned long long v[16];
        tile_prof_read(v);
        for (int i = 0; i < 16; i++)
            out16[i] = v[i];
    });
}
int b2sv_comm_stats(const b2sv_state *s, uint64_t *swaps, uint64_t *swap_bytes, int *peer_path) {
    return guard([&] {
        uint64_t a = 0, b = 0;
        int p = 0;
        st(s).comm_stats(&a, &b, &p);
        if (swaps)
            *swaps = a;
        if (swap_bytes)
            *swap_bytes = b;
        if (peer_path)
            *peer_path = p;
    });
}
int b2sv_last_upload_bytes(const b2sv_state *s, uint64_t *bytes) {
    return guard([&] { *bytes = st(s).last_upload_bytes(); });
}
int b2sv_layout(const b2sv_state *s, int *l2p, int cap, int *n) {
    return guard([&] {
        B2_ABORT_IF(!n, "null output");
        *n = st(s).num_qubits();
        for (int q = 0; q < *n && q < cap; q++)
            l2p[q] = st(s).phys_bit(q);
    });
}
int b2sv_normalize_layout(b2sv_state *s) {
    return guard([&] { st(s).normalize_layout(); });
}

// Host-only: how the fusion scheduler would execute `ops` on an n-qubit state (no device needed).
int b2sv_plan_ops(const b2sv_ops *ops, int num_qubits, int dtype, uint64_t *passes,
                  uint64_t *rounds, uint64_t *arithmetic_ops, uint64_t *absorbed_perms,
                  uint64_t *fused_stores) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null op list");
        B2_ABORT_IF(dtype != 0 && dtype != 1, "dtype must be B2SV_C64 (0) or B2SV_C128 (1)");
        std::vector<Prim> prims;
        for (const GateOp &op : ops->d.ops) {
            if (op.name == "Identity")
                continue;
            const std::vector<int> bits = wires_to_bits(op.wires, num_qubits);
            if (!lower_gate(op.name, bits, op.inverse, op.params, prims)) {
                B2_ABORT_IF(op.matrix.empty(), "operation '" + op.name +
                                                   "' is not a named gate and no matrix was provided");
                lower_matrix(bits, op.inverse, op.matrix, prims);
            }
        }
        SchedConfig cfg;
        tile_config(dtype, &cfg.B, &cfg.R);
        cfg.SW = dtype == 1 ? 3 : 4;
        cfg.SH = dtype == 1 ? 0 : 1;
        cfg.f32 = dtype != 1;
        cfg.n_local = num_qubits;
        cfg.n_alloc = std::max(num_qubits, cfg.B);
        cfg.low = default_tile_low();
        uint64_t np = 0, nr = 0, na = 0, nabs = 0, nf = 0;
        if (!prims.empty())
            for (const Pass &ps : build_schedule(prims, cfg)) {
                np++;
                if (ps.is_matk) {
                    na++;
                    continue;
                }
                nr += ps.hdr.n_rounds;
                na += ps.hdr.n_ops;
                nabs += ps.n_absorbed;
                nf += ps.hdr.fused_store != 0;
            }
        if (passes)
            *passes = np;
        if (rounds)
            *rounds = nr;
        if (arithmetic_ops)
            *arithmetic_ops = na;
        if (absorbed_perms)
            *absorbed_perms = nabs;
        if (fused_stores)
            *fused_stores = nf;
    });
}

// Host-only: how `ops` would run on a state sharded over `world` ranks -- the runs and exchanges of
// shard_plan.cpp, the tile passes of every run. stats: [0] runs, [1] exchanges, [2] tile passes,
// [3] exchanged bits in total, [4] bytes each rank sends (complex128: 16 B per amplitude).
// If buf != NULL the plan is written as text (one line per step / primitive) for emulation in tests,
// followed by a line "L2P ..." with the final logical -> physical bit map.
int b2sv_plan_sharded(const b2sv_ops *ops, int num_qubits, int world, int dtype, uint64_t *stats5,
                      char *buf, size_t cap) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null op list");
        int g = 0;
        while ((1 << g) < world)
            g++;
        B2_ABORT_IF((1 << g) != world || num_qubits - g < 1, "invalid world size");
        std::vector<Prim> prims;
        for (const GateOp &op : ops->d.ops) {
            if (op.name == "Identity")
                continue;
            const std::vector<int> bits = wires_to_bits(op.wires, num_qubits);
            if (!lower_gate(op.name, bits, op.inverse, op.params, prims)) {
                B2_ABORT_IF(op.matrix.empty(), "operation '" + op.name +
                                                   "' is not a named gate and no matrix was provided");
                lower_matrix(bits, op.inverse, op.matrix, prims);
            }
        }
        ShardPlanConfig pc;
        pc.n = num_qubits;
        pc.n_local = num_qubits - g;
        pc.min_victim_pos = std::min(5, std::max(0, pc.n_local - g - 1));
        std::vector<int> l2p(num_qubits);
        for (int q = 0; q < num_qubits; q++)
            l2p[q] = q;
        const std::vector<ShardStep> steps = plan_sharded(prims, l2p, pc);
        SchedConfig cfg;
        tile_config(dtype, &cfg.B, &cfg.R);
        cfg.SW = dtype == 1 ? 3 : 4;
        cfg.SH = dtype == 1 ? 0 : 1;
        cfg.f32 = dtype != 1;
        cfg.n_local = pc.n_local;
        cfg.n_alloc = std::max(pc.n_local, cfg.B);
        cfg.low = default_tile_low();
        uint64_t st[5] = {0, 0, 0, 0, 0};
        std::ostringstream os;
        os << std::setprecision(17);
        for (const ShardStep &sp : steps) {
            if (sp.is_exchange) {
                st[1]++;
                st[3] += sp.swaps.size();
                st[4] += ((uint64_t(1) << pc.n_local) - (uint64_t(1) << (pc.n_local - sp.swaps.size()))) * 16;
                if (buf) {
                    os << "EXCH";
                    for (const auto &pr : sp.swaps)
                        os << ' ' << pr.first << ' ' << pr.second;
                    os << '\n';
                }
                continue;
            }
            st[0]++;
            st[2] += build_schedule(sp.prims, cfg).size();
            if (!buf)
                continue;
            os << "RUN " << sp.prims.size() << '\n';
            for (const Prim &p : sp.prims) {
                if (p.type == Prim::C1Q) {
                    os << "C1Q " << p.target << ' ' << p.cmask << ' ' << p.cval;
                    for (int i = 0; i < 4; i++)
                        os << ' ' << p.m[i].real() << ' ' << p.m[i].imag();
                } else if (p.type == Prim::DIAG) {
                    os << "DIAG " << p.pmask << ' ' << p.cmask << ' ' << p.cval;
                    for (int i = 0; i < 2; i++)
                        os << ' ' << p.m[i].real() << ' ' << p.m[i].imag();
                } else {
                    os << "MATK " << p.bits.size();
                    for (int b : p.bits)
                        os << ' ' << b;
                    for (const cplx &c : p.mat)
                        os << ' ' << c.real() << ' ' << c.imag();
                }
                os << '\n';
            }
        }
        if (buf) {
            os << "L2P";
            for (int q = 0; q < num_qubits; q++)
                os << ' ' << l2p[q];
            os << '\n';
            const std::string str = os.str();
            B2_ABORT_IF(str.size() + 1 > cap, "plan text does not fit the buffer");
            std::memcpy(buf, str.c_str(), str.size() + 1);
        }
        if (stats5)
            for (int i = 0; i < 5; i++)
                stats5[i] = st[i];
    });
}

int b2sv_ops_create(int nops, const char *const *names, const double *params, const int *nparams,
                    const int64_t *wires, const int *nwires, const int *inverses,
                    const double *const *matrices, b2sv_ops **out) {
    return guard([&] {
        B2_ABORT_IF(nops < 0 || !out, "invalid arguments");
        auto h = std::make_unique<b2sv_ops>();
        size_t po = 0, wo = 0;
        for (int i = 0; i < nops; i++) {
            GateOp op;
            B2_ABORT_IF(!names || !nparams || !nwires || !inverses, "null argument");
            op.name = name_str(names[i]);
            B2_ABORT_IF(nparams[i] < 0 || nwires[i] < 0, "negative parameter or wire count");
            B2_ABORT_IF((nparams[i] > 0 && !params) || (nwires[i] > 0 && !wires), "null argument");
            op.params.assign(params + po, params + po + nparams[i]);
            po += nparams[i];
            op.wires.assign(wires + wo, wires + wo + nwires[i]);
            wo += nwires[i];
            op.inverse = inverses[i] != 0;
            // named gates never use their matrix (reference StateVectorKokkos.hpp:585-600)
            if (matrices && matrices[i] && op.name != "Identity" && !is_named_gate(op.name))
                op.matrix = cplx_vec(matrices[i], matrix_len(nwires[i]));
            if (!op.params.empty())
                h->d.num_par_ops++;
            h->d.ops.push_back(std::move(op));
        }
        *out = h.release();
    });
}
int b2sv_ops_destroy(b2sv_ops *ops) {
    return guard([&] { delete ops; });
}
int b2sv_ops_size(const b2sv_ops *ops, int *nops, int *n_par_ops) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null ops handle");
        if (nops)
            *nops = static_cast<int>(ops->d.ops.size());
        if (n_par_ops)
            *n_par_ops = static_cast<int>(ops->d.num_par_ops);
    });
}

int b2sv_expval_named(const b2sv_state *s, const char *name, const int64_t *wires, int nw,
                      double *out) {
    return guard([&] { *out = st(s).expval_named(name_str(name), wires_vec(wires, nw)); });
}
int b2sv_expval_matrix(const b2sv_state *s, const int64_t *wires, int nw, const double *matrix,
                       double *out) {
    return guard([&] {
        *out = st(s).expval_matrix(wires_vec(wires, nw), cplx_vec(matrix, matrix_len(nw)));
    });
}
int b2sv_expval_csr(const b2sv_state *s, const double *data, const uint64_t *indices,
                    const uint64_t *indptr, size_t nnz, size_t nrows, double *out) {
    return guard([&] {
        auto d = cplx_vec(data, nnz);
        auto m = csr_upload(st(s).device(), d.data(), indices, indptr, nnz, nrows);
        *out = st(s).expval_csr(*m);
    });
}
int b2sv_csr_create(const b2sv_state *like, const double *data, const uint64_t *indices,
                    const uint64_t *indptr, size_t nnz, size_t nrows, b2sv_csr **out) {
    return guard([&] {
        auto d = cplx_vec(data, nnz);
        auto h = std::make_unique<b2sv_csr>();
        h->m = csr_upload(st(like).device(), d.data(), indices, indptr, nnz, nrows);
        *out = h.release();
    });
}
int b2sv_csr_destroy(b2sv_csr *m) {
    return guard([&] { delete m; });
}
int b2sv_expval_csr_resident(const b2sv_state *s, const b2sv_csr *m, double *out) {
    return guard([&] {
        B2_ABORT_IF(!m || !m->m, "null CSR handle");
        *out = st(s).expval_csr(*m->m);
    });
}
int b2sv_expval_obs(const b2sv_state *s, const b2sv_obs *ob, double *out) {
    return guard([&] {
        B2_ABORT_IF(!ob || !ob->o, "null observable handle");
        *out = expval_obs(st(s), *ob->o);
    });
}
int b2sv_var_obs(const b2sv_state *s, const b2sv_obs *ob, double *out) {
    return guard([&] {
        B2_ABORT_IF(!ob || !ob->o, "null observable handle");
        *out = var_obs(st(s), *ob->o);
    });
}
int b2sv_probs(const b2sv_state *s, const int64_t *wires, int nw, double *out) {
    return guard([&] { st(s).probs(wires_vec(wires, nw), out); });
}
int b2sv_generate_samples(const b2sv_state *s, size_t shots, uint64_t seed, uint64_t *out) {
    return guard([&] { st(s).generate_samples(shots, seed, out); });
}
int b2sv_inner_product(const b2sv_state *a, const b2sv_state *b, double *re, double *im) {
    return guard([&] { st(a).inner_product(st(b), re, im); });
}
int b2sv_axpy(double ar, double ai, const b2sv_state *x, b2sv_state *y) {
    return guard([&] { st(y).axpy(cplx(ar, ai), st(x)); });
}

int b2sv_obs_named(const char *name, const int64_t *wires, int nw, b2sv_obs **out) {
    return guard([&] { *out = new b2sv_obs{make_named_obs(name_str(name), wires_vec(wires, nw))}; });
}
int b2sv_obs_hermitian(const double *matrix, const int64_t *wires, int nw, b2sv_obs **out) {
    return guard([&] {
        *out = new b2sv_obs{
            make_hermitian_obs(cplx_vec(matrix, matrix_len(nw)), wires_vec(wires, nw))};
    });
}
int b2sv_obs_tensor(b2sv_obs *const *obs, int n, b2sv_obs **out) {
    return guard([&] {
        std::vector<ObsPtr> v;
        for (int i = 0; i < n; i++) {
            B2_ABORT_IF(!obs[i] || !obs[i]->o, "null observable handle");
            v.push_back(obs[i]->o);
        }
        *out = new b2sv_obs{make_tensor_obs(v)};
    });
}
int b2sv_obs_hamiltonian(const double *coeffs, b2sv_obs *const *obs, int n, b2sv_obs **out) {
    return guard([&] {
        std::vector<ObsPtr> v;
        for (int i = 0; i < n; i++) {
            B2_ABORT_IF(!obs[i] || !obs[i]->o, "null observable handle");
            v.push_back(obs[i]->o);
        }
        *out = new b2sv_obs{make_hamiltonian_obs(std::vector<double>(coeffs, coeffs + n), v)};
    });
}
int b2sv_obs_sparse(const double *data, const uint64_t *indices, const uint64_t *indptr, size_t nnz,
                    size_t nrows, const int64_t *wires, int nw, b2sv_obs **out) {
    return guard([&] {
        *out = new b2sv_obs{make_sparse_obs(cplx_vec(data, nnz),
                                            std::vector<uint64_t>(indices, indices + nnz),
                                            std::vector<uint64_t>(indptr, indptr + nrows + 1),
                                            wires_vec(wires, nw))};
    });
}
int b2sv_obs_destroy(b2sv_obs *ob) {
    return guard([&] { delete ob; });
}
int b2sv_obs_name(const b2sv_obs *ob, char *buf, size_t cap) {
    return guard([&] {
        B2_ABORT_IF(!ob || !ob->o, "null observable handle");
        copy_str(ob->o->name(), buf, cap);
    });
}
int b2sv_obs_wires(const b2sv_obs *ob, int64_t *wires, int cap, int *nw) {
    return guard([&] {
        B2_ABORT_IF(!ob || !ob->o, "null observable handle");
        const auto w = ob->o->wires();
        *nw = static_cast<int>(w.size());
        for (int i = 0; i < cap && i < *nw; i++)
            wires[i] = w[i];
    });
}
int b2sv_obs_apply(const b2sv_obs *ob, b2sv_state *s) {
    return guard([&] {
        B2_ABORT_IF(!ob || !ob->o, "null observable handle");
        ob->o->apply_in_place(st(s));
    });
}

int b2sv_adjoint_jacobian(const b2sv_state *s, b2sv_obs *const *obs, int n_obs,
                          const b2sv_ops *ops, const uint64_t *trainable_params, int n_tp,
                          double *jac_out) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null ops handle");
        std::vector<ObsPtr> v;
        for (int i = 0; i < n_obs; i++) {
            B2_ABORT_IF(!obs[i] || !obs[i]->o, "null observable handle");
            v.push_back(obs[i]->o);
        }
        adjoint_jacobian(st(s), v, ops->d,
                         std::vector<uint64_t>(trainable_params, trainable_params + n_tp), jac_out);
    });
}
int b2sv_adjoint_vjp(const b2sv_state *s, b2sv_obs *const *obs, int n_obs, const double *dy,
                     const b2sv_ops *ops, const uint64_t *trainable_params, int n_tp,
                     double *vjp_out) {
    return guard([&] {
        B2_ABORT_IF(!ops, "null ops handle");
        B2_ABORT_IF(!dy || !vjp_out, "null buffer");
        // lightning_kokkos.py:689-727: the vector-Jacobian product of expectation values is the
        // adjoint Jacobian of ONE observable, the Hamiltonian sum_i dy_i O_i -- one reverse sweep
        // whatever the number of measurements
        std::vector<ObsPtr> v;
        for (int i = 0; i < n_obs; i++) {
            B2_ABORT_IF(!obs[i] || !obs[i]->o, "null observable handle");
            v.push_back(obs[i]->o);
        }
        bool all_zero = true;
        for (int i = 0; i < n_obs; i++)
            all_zero = all_zero && dy[i] == 0.0;
        if (all_zero || n_tp == 0) { // lightning_kokkos.py:701-702
            for (int i = 0; i < n_tp; i++)
                vjp_out[i] = 0.0;
            return;
        }
        std::vector<ObsPtr> ham = {make_hamiltonian_obs(std::vector<double>(dy, dy + n_obs), v)};
        adjoint_jacobian(st(s), ham, ops->d,
                         std::vector<uint64_t>(trainable_params, trainable_params + n_tp), vjp_out);
    });
}
}
