// TEST INFRASTRUCTURE ONLY -- not part of the product, never linked by it.
//
// Thin extern "C" driver over the UNMODIFIED reference headers
// (/root/reference/pennylane_lightning_kokkos/src/{simulator,util,algorithms}) compiled
// against oracle/kokkos_shim. It exposes the reference's own StateVectorKokkos /
// MeasuresKokkos / ObservablesKokkos / AdjointJacobianKokkos so that tests can pin the
// NumPy restatement (oracle/np_oracle.py) and the CUDA engine against the reference's
// per-amplitude arithmetic, and so that bench.py can time the reference's CPU path
// (Kokkos-OpenMP semantics) on the GPU box's host cores.
//
// Built by oracle/Makefile into oracle/_ref/libref_oracle.so (git-ignored, travels
// with gpurun). No reference source is copied: everything is reached by -I.
#include <complex>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "AdjointDiffKokkos.hpp"
#include "MeasuresKokkos.hpp"
#include "ObservablesKokkos.hpp"
#include "StateVectorKokkos.hpp"

using namespace Pennylane;
using namespace Pennylane::Lightning_Kokkos::Simulators;
using namespace Pennylane::Lightning_Kokkos::Algorithms;

namespace {
thread_local std::string g_err;

template <class P> struct SV {
    StateVectorKokkos<P> sv;
    explicit SV(size_t n) : sv(n) {}
};
struct Handle {
    int prec; // 0 = float, 1 = double
    void *p;
};
template <class P> StateVectorKokkos<P> &sv_of(void *h) {
    return static_cast<SV<P> *>(static_cast<Handle *>(h)->p)->sv;
}
struct ObsHandle {
    int prec;
    std::shared_ptr<ObservableKokkos<float>> f;
    std::shared_ptr<ObservableKokkos<double>> d;
};
template <class P> std::shared_ptr<ObservableKokkos<P>> &obs_of(ObsHandle *o) {
    if constexpr (std::is_same_v<P, float>)
        return o->f;
    else
        return o->d;
}
template <class P> std::vector<P> to_prec(const double *p, int n) {
    std::vector<P> v(n);
    for (int i = 0; i < n; i++)
        v[i] = static_cast<P>(p[i]);
    return v;
}
template <class C, class P> std::vector<C> to_cplx(const double *p, size_t n) {
    std::vector<C> v(n);
    for (size_t i = 0; i < n; i++)
        v[i] = C(static_cast<P>(p[2 * i]), static_cast<P>(p[2 * i + 1]));
    return v;
}
std::vector<size_t> to_wires(const int64_t *w, int n) {
    return std::vector<size_t>(w, w + n);
}

#define GUARD(body)                                                                         \
    try {                                                                                   \
        body;                                                                               \
        return 0;                                                                           \
    } catch (const std::exception &e) {                                                     \
        g_err = e.what();                                                                   \
        return 1;                                                                           \
    }
#define DISPATCH(h, fn, ...)                                                                \
    (static_cast<Handle *>(h)->prec == 0 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

template <class P> void apply_named(void *h, const char *name, const int64_t *w, int nw, int inv,
                                    const double *par, int np) {
    sv_of<P>(h).applyOperation(std::string(name), to_wires(w, nw), inv != 0,
                               to_prec<P>(par, np));
}
template <class P>
void apply_matrix(void *h, const int64_t *w, int nw, int inv, const double *mat) {
    const size_t dim = size_t(1) << nw;
    auto m = to_cplx<Kokkos::complex<P>, P>(mat, dim * dim);
    sv_of<P>(h).applyOperation_std("__matrix__", to_wires(w, nw), inv != 0, {}, m);
}
template <class P>
double apply_generator(void *h, const char *name, const int64_t *w, int nw, int adj) {
    return static_cast<double>(
        sv_of<P>(h).applyGenerator(std::string(name), to_wires(w, nw), adj != 0));
}
template <class P> double expval_named(void *h, const char *name, const int64_t *w, int nw) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    return static_cast<double>(m.getExpectationValue(std::string(name), to_wires(w, nw)));
}
template <class P> double expval_matrix(void *h, const int64_t *w, int nw, const double *mat) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    const size_t dim = size_t(1) << nw;
    return static_cast<double>(
        m.getExpectationValue(to_wires(w, nw), to_cplx<Kokkos::complex<P>, P>(mat, dim * dim)));
}
template <class P>
double expval_csr(void *h, const double *data, const int64_t *indices, const int64_t *indptr,
                  int64_t nnz, int64_t nrows) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    return static_cast<double>(m.getExpectationValue(
        to_cplx<Kokkos::complex<P>, P>(data, nnz), std::vector<size_t>(indices, indices + nnz),
        std::vector<size_t>(indptr, indptr + nrows + 1)));
}
template <class P> void probs(void *h, const int64_t *w, int nw, int all, double *out) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    std::vector<P> p = all ? m.probs() : m.probs(to_wires(w, nw));
    for (size_t i = 0; i < p.size(); i++)
        out[i] = static_cast<double>(p[i]);
}
template <class P> void samples(void *h, int64_t shots, uint64_t *out) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    auto s = m.generate_samples(static_cast<size_t>(shots));
    for (size_t i = 0; i < s.size(); i++)
        out[i] = s[i];
}
template <class P> double expval_obs(void *h, ObsHandle *o) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    return static_cast<double>(m.expval(*obs_of<P>(o)));
}
template <class P> double var_obs(void *h, ObsHandle *o) {
    MeasuresKokkos<P> m(sv_of<P>(h));
    return static_cast<double>(m.var(*obs_of<P>(o)));
}
template <class P> void obs_apply(void *h, ObsHandle *o) { obs_of<P>(o)->applyInPlace(sv_of<P>(h)); }

template <class P>
void adjoint(void *h, ObsHandle **obs, int nobs, int nops, const char **names,
             const double *params, const int *nparams, const int64_t *wires, const int *nwires,
             const int *inverses, const int64_t *tp, int ntp, double *jac_out) {
    std::vector<std::string> ops_name(names, names + nops);
    std::vector<std::vector<P>> ops_params(nops);
    std::vector<std::vector<size_t>> ops_wires(nops);
    std::vector<bool> ops_inv(nops);
    size_t po = 0, wo = 0;
    for (int i = 0; i < nops; i++) {
        for (int k = 0; k < nparams[i]; k++)
            ops_params[i].push_back(static_cast<P>(params[po++]));
        for (int k = 0; k < nwires[i]; k++)
            ops_wires[i].push_back(static_cast<size_t>(wires[wo++]));
        ops_inv[i] = inverses[i] != 0;
    }
    AdjointJacobianKokkos<P> adj;
    auto ops = adj.createOpsData(ops_name, ops_params, ops_wires, ops_inv,
                                 std::vector<std::vector<std::complex<P>>>(nops));
    std::vector<std::shared_ptr<ObservableKokkos<P>>> ovec;
    for (int i = 0; i < nobs; i++)
        ovec.push_back(obs_of<P>(obs[i]));
    std::vector<std::vector<P>> jac(nobs, std::vector<P>(ntp, 0));
    adj.adjointJacobian(sv_of<P>(h), jac, ovec, ops, std::vector<size_t>(tp, tp + ntp), false);
    for (int o = 0; o < nobs; o++)
        for (int p = 0; p < ntp; p++)
            jac_out[size_t(o) * ntp + p] = static_cast<double>(jac[o][p]);
}
} // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void *ref_sv_create(int prec, int num_qubits) {
    try {
        auto *h = new Handle{prec, nullptr};
        if (prec == 0)
            h->p = new SV<float>(num_qubits);
        else
            h->p = new SV<double>(num_qubits);
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void ref_sv_destroy(void *h) {
    auto *hh = static_cast<Handle *>(h);
    if (hh->prec == 0)
        delete static_cast<SV<float> *>(hh->p);
    else
        delete static_cast<SV<double> *>(hh->p);
    delete hh;
}
int ref_sv_reset(void *h) {
    GUARD(if (static_cast<Handle *>(h)->prec == 0) sv_of<float>(h).resetStateVector();
          else sv_of<double>(h).resetStateVector())
}
int ref_sv_set_basis_state(void *h, int64_t index) {
    GUARD(if (static_cast<Handle *>(h)->prec == 0) sv_of<float>(h).setBasisState(index);
          else sv_of<double>(h).setBasisState(index))
}
int ref_sv_set_state_vector(void *h, const int64_t *indices, const double *values, int64_t n) {
    GUARD({
        std::vector<size_t> idx(indices, indices + n);
        if (static_cast<Handle *>(h)->prec == 0)
            sv_of<float>(h).setStateVector(idx, to_cplx<Kokkos::complex<float>, float>(values, n));
        else
            sv_of<double>(h).setStateVector(idx,
                                            to_cplx<Kokkos::complex<double>, double>(values, n));
    })
}
// host buffers are in the state's own precision (complex64 / complex128), interleaved
int ref_sv_h2d(void *h, void *host, int64_t length) {
    GUARD(if (static_cast<Handle *>(h)->prec == 0)
              sv_of<float>(h).HostToDevice(static_cast<Kokkos::complex<float> *>(host), length);
          else sv_of<double>(h).HostToDevice(static_cast<Kokkos::complex<double> *>(host), length))
}
int ref_sv_d2h(void *h, void *host, int64_t length) {
    GUARD(if (static_cast<Handle *>(h)->prec == 0)
              sv_of<float>(h).DeviceToHost(static_cast<Kokkos::complex<float> *>(host), length);
          else sv_of<double>(h).DeviceToHost(static_cast<Kokkos::complex<double> *>(host), length))
}
// sampled read: out[2k], out[2k+1] = amplitude at indices[k] (parity checks on states too big to copy)
int ref_sv_get_amplitudes(void *h, const int64_t *indices, int64_t n, double *out) {
    GUARD(for (int64_t k = 0; k < n; k++) {
        if (static_cast<Handle *>(h)->prec == 0) {
            const auto v = sv_of<float>(h).getData().data()[indices[k]];
            out[2 * k] = v.real();
            out[2 * k + 1] = v.imag();
        } else {
            const auto v = sv_of<double>(h).getData().data()[indices[k]];
            out[2 * k] = v.real();
            out[2 * k + 1] = v.imag();
        }
    })
}
int ref_sv_apply(void *h, const char *name, const int64_t *wires, int nw, int inverse,
                 const double *params, int np) {
    GUARD(DISPATCH(h, apply_named, h, name, wires, nw, inverse, params, np))
}
int ref_sv_apply_matrix(void *h, const int64_t *wires, int nw, int inverse, const double *mat) {
    GUARD(DISPATCH(h, apply_matrix, h, wires, nw, inverse, mat))
}
int ref_sv_apply_generator(void *h, const char *name, const int64_t *wires, int nw, int adj,
                           double *scale) {
    GUARD(*scale = DISPATCH(h, apply_generator, h, name, wires, nw, adj))
}
int ref_expval_named(void *h, const char *name, const int64_t *wires, int nw, double *out) {
    GUARD(*out = DISPATCH(h, expval_named, h, name, wires, nw))
}
int ref_expval_matrix(void *h, const int64_t *wires, int nw, const double *mat, double *out) {
    GUARD(*out = DISPATCH(h, expval_matrix, h, wires, nw, mat))
}
int ref_expval_csr(void *h, const double *data, const int64_t *indices, const int64_t *indptr,
                   int64_t nnz, int64_t nrows, double *out) {
    GUARD(*out = DISPATCH(h, expval_csr, h, data, indices, indptr, nnz, nrows))
}
int ref_probs(void *h, const int64_t *wires, int nw, int all, double *out) {
    GUARD(DISPATCH(h, probs, h, wires, nw, all, out))
}
int ref_generate_samples(void *h, int64_t shots, uint64_t *out) {
    GUARD(DISPATCH(h, samples, h, shots, out))
}

// ---- observables -------------------------------------------------------------------
void *ref_obs_named(int prec, const char *name, const int64_t *wires, int nw) {
    auto *o = new ObsHandle{prec, nullptr, nullptr};
    if (prec == 0)
        o->f = std::make_shared<NamedObsKokkos<float>>(std::string(name), to_wires(wires, nw));
    else
        o->d = std::make_shared<NamedObsKokkos<double>>(std::string(name), to_wires(wires, nw));
    return o;
}
void *ref_obs_hermitian(int prec, const double *mat, const int64_t *wires, int nw) {
    auto *o = new ObsHandle{prec, nullptr, nullptr};
    const size_t dim = size_t(1) << nw;
    if (prec == 0)
        o->f = std::make_shared<HermitianObsKokkos<float>>(
            to_cplx<std::complex<float>, float>(mat, dim * dim), to_wires(wires, nw));
    else
        o->d = std::make_shared<HermitianObsKokkos<double>>(
            to_cplx<std::complex<double>, double>(mat, dim * dim), to_wires(wires, nw));
    return o;
}
void *ref_obs_tensor(int prec, void **obs, int n) {
    try {
        auto *o = new ObsHandle{prec, nullptr, nullptr};
        if (prec == 0) {
            std::vector<std::shared_ptr<ObservableKokkos<float>>> v;
            for (int i = 0; i < n; i++)
                v.push_back(static_cast<ObsHandle *>(obs[i])->f);
            o->f = TensorProdObsKokkos<float>::create(v);
        } else {
            std::vector<std::shared_ptr<ObservableKokkos<double>>> v;
            for (int i = 0; i < n; i++)
                v.push_back(static_cast<ObsHandle *>(obs[i])->d);
            o->d = TensorProdObsKokkos<double>::create(v);
        }
        return o;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *ref_obs_hamiltonian(int prec, const double *coeffs, void **obs, int n) {
    auto *o = new ObsHandle{prec, nullptr, nullptr};
    if (prec == 0) {
        std::vector<std::shared_ptr<ObservableKokkos<float>>> v;
        for (int i = 0; i < n; i++)
            v.push_back(static_cast<ObsHandle *>(obs[i])->f);
        o->f = std::make_shared<HamiltonianKokkos<float>>(to_prec<float>(coeffs, n), v);
    } else {
        std::vector<std::shared_ptr<ObservableKokkos<double>>> v;
        for (int i = 0; i < n; i++)
            v.push_back(static_cast<ObsHandle *>(obs[i])->d);
        o->d = std::make_shared<HamiltonianKokkos<double>>(to_prec<double>(coeffs, n), v);
    }
    return o;
}
void *ref_obs_sparse(int prec, const double *data, const int64_t *indices, const int64_t *indptr,
                     int64_t nnz, int64_t nrows, const int64_t *wires, int nw) {
    auto *o = new ObsHandle{prec, nullptr, nullptr};
    std::vector<size_t> ind(indices, indices + nnz), ptr(indptr, indptr + nrows + 1);
    if (prec == 0)
        o->f = std::make_shared<SparseHamiltonianKokkos<float>>(
            to_cplx<std::complex<float>, float>(data, nnz), ind, ptr, to_wires(wires, nw));
    else
        o->d = std::make_shared<SparseHamiltonianKokkos<double>>(
            to_cplx<std::complex<double>, double>(data, nnz), ind, ptr, to_wires(wires, nw));
    return o;
}
void ref_obs_destroy(void *o) { delete static_cast<ObsHandle *>(o); }
int ref_obs_name(void *o, char *buf, int cap) {
    GUARD({
        auto *oh = static_cast<ObsHandle *>(o);
        std::string s = oh->prec == 0 ? oh->f->getObsName() : oh->d->getObsName();
        std::strncpy(buf, s.c_str(), cap - 1);
        buf[cap - 1] = 0;
    })
}
int ref_expval_obs(void *h, void *o, double *out) {
    GUARD(*out = DISPATCH(h, expval_obs, h, static_cast<ObsHandle *>(o)))
}
int ref_var_obs(void *h, void *o, double *out) {
    GUARD(*out = DISPATCH(h, var_obs, h, static_cast<ObsHandle *>(o)))
}
int ref_obs_apply(void *h, void *o) { GUARD(DISPATCH(h, obs_apply, h, static_cast<ObsHandle *>(o))) }

// ---- adjoint Jacobian --------------------------------------------------------------
// ops are flattened: params / wires are concatenated, nparams[i] / nwires[i] give the split.
int ref_adjoint_jacobian(void *h, void **obs, int nobs, int nops, const char **names,
                         const double *params, const int *nparams, const int64_t *wires,
                         const int *nwires, const int *inverses, const int64_t *tp, int ntp,
                         double *jac_out) {
    GUARD(DISPATCH(h, adjoint, h, reinterpret_cast<ObsHandle **>(obs), nobs, nops, names, params,
                   nparams, wires, nwires, inverses, tp, ntp, jac_out))
}
}
