// b2sv: compiled pybind11 module with the reference's binding surface
// (reference pennylane_lightning_kokkos/src/bindings/Bindings.cpp:50-968) over the C ABI (include/b2sv.h).
// Same class names, method names, argument order and exception type as the reference module
// `lightning_kokkos_qubit_ops`; it links libb2sv.so instead of Kokkos. The ctypes mirror
// (lightning_kokkos_qubit_ops.py) exposes the same surface without a compile step.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <complex>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "b2sv.h"

namespace py = pybind11;
using cplx = std::complex<double>;

namespace {
struct PLError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
void chk(int rc) {
    if (rc)
        throw PLError(b2sv_last_error());
}

const char *kGates[] = {"Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "T", "CNOT", "SWAP",
                        "CSWAP", "Toffoli", "CY", "CZ", "PhaseShift", "ControlledPhaseShift", "RX", "RY",
                        "RZ", "Rot", "CRX", "CRY", "CRZ", "CRot", "IsingXX", "IsingXY", "IsingYY",
                        "IsingZZ", "MultiRZ", "SingleExcitation", "SingleExcitationMinus",
                        "SingleExcitationPlus", "DoubleExcitation", "DoubleExcitationMinus",
                        "DoubleExcitationPlus"};
bool is_gate(const std::string &n) {
    for (const char *g : kGates)
        if (n == g)
            return true;
    return false;
}

// Inert mirror of Kokkos::InitializationSettings (Bindings.cpp:854-967); only device_id is honoured.
struct InitSettings {
    int num_threads = 0, device_id = 0;
    std::string map_device_id_by, tools_libs, tools_args;
    bool disable_warnings = false, print_configuration = false, tune_internals = false, tools_help = false;
    unsigned has = 0x1ffu; // the reference's py::init calls every setter (Bindings.cpp:855-866)
};

template <int DT> struct ObsT { // ObservableKokkos<P> hierarchy (Bindings.cpp:591-736)
    b2sv_obs *h = nullptr;
    std::vector<std::shared_ptr<ObsT<DT>>> children;
    ~ObsT() { b2sv_obs_destroy(h); }
    std::string name() const {
        std::string buf(1 << 20, '\0');
        chk(b2sv_obs_name(h, buf.data(), buf.size()));
        buf.resize(std::strlen(buf.c_str()));
        return buf;
    }
    std::vector<int64_t> wires() const {
        int n = 0;
        std::vector<int64_t> w(64);
        chk(b2sv_obs_wires(h, w.data(), 64, &n));
        w.resize(n);
        return w;
    }
};

template <int DT> struct OpsT { // OpsData<P> (AdjointDiffKokkos.hpp:17-173)
    b2sv_ops *h = nullptr;
    std::vector<std::string> names;
    std::vector<std::vector<double>> params;
    std::vector<bool> inverses;
    ~OpsT() { b2sv_ops_destroy(h); }
};
template <int DT> struct AdjT {};

template <int DT>
std::shared_ptr<OpsT<DT>> make_ops(const std::vector<std::string> &names,
                              const std::vector<std::vector<double>> &params,
                              const std::vector<std::vector<int64_t>> &wires, const std::vector<bool> &inv,
                              const std::vector<std::vector<cplx>> &mats) {
    const size_t n = names.size();
    if (params.size() != n || wires.size() != n || inv.size() != n)
        throw PLError("Incompatible number of ops, params, wires and inverses");
    auto o = std::make_shared<OpsT<DT>>();
    o->names = names;
    o->params = params;
    o->inverses = inv;
    std::vector<const char *> cn(n);
    std::vector<double> fp;
    std::vector<int64_t> fw;
    std::vector<int> np(n), nw(n), iv(n);
    std::vector<const double *> mp(n, nullptr);
    for (size_t i = 0; i < n; i++) {
        cn[i] = names[i].c_str();
        np[i] = static_cast<int>(params[i].size());
        nw[i] = static_cast<int>(wires[i].size());
        iv[i] = inv[i] ? 1 : 0;
        fp.insert(fp.end(), params[i].begin(), params[i].end());
        fw.insert(fw.end(), wires[i].begin(), wires[i].end());
        if (i < mats.size() && !mats[i].empty()) {
            if (mats[i].size() != (size_t(1) << (2 * wires[i].size())))
                throw PLError("matrix size does not match the number of wires");
            mp[i] = reinterpret_cast<const double *>(mats[i].data());
        }
    }
    chk(b2sv_ops_create(static_cast<int>(n), cn.data(), fp.data(), np.data(), fw.data(), nw.data(), iv.data(),
                        mp.data(), &o->h));
    return o;
}

template <int DT> struct SV { // StateVectorKokkos<P> + MeasuresKokkos<P> (Bindings.cpp:59-585)
    b2sv_state *h = nullptr;
    explicit SV(int n, int dev = 0) { chk(b2sv_create(n, DT, dev, &h)); }
    ~SV() { b2sv_destroy(h); }
    SV(const SV &) = delete;
    int nq() const {
        int n = 0;
        chk(b2sv_num_qubits(h, &n));
        return n;
    }
};

template <int DT, class np_c, class np_r> void bind_precision(py::module_ &m, const std::string &sfx) {
    using S = SV<DT>;
    using Obs = ObsT<DT>;
    using ObsP = std::shared_ptr<Obs>;
    using Ops = OpsT<DT>;
    using Adj = AdjT<DT>;
    using arr_c = py::array_t<np_c, py::array::c_style | py::array::forcecast>;
    auto cls = py::class_<S>(m, ("LightningKokkos_" + sfx).c_str());
    cls.def(py::init([](int n) { return new S(n); }))                                   // :64-66
        .def(py::init([](int n, const InitSettings &s) { return new S(n, s.device_id); })) // :67-70
        .def(py::init([](const arr_c &a) {                                                // :71-77
            size_t len = static_cast<size_t>(a.size());
            int n = 0;
            while ((size_t(1) << n) < len)
                n++;
            if (len == 0 || (size_t(1) << n) != len)
                throw PLError("state vector length must be a power of two");
            auto *s = new S(n);
            chk(b2sv_h2d(s->h, a.data(), len));
            return s;
        }))
        .def(py::init([](const arr_c &a, const InitSettings &st) {                        // :78-85
            size_t len = static_cast<size_t>(a.size());
            int n = 0;
            while ((size_t(1) << n) < len)
                n++;
            if (len == 0 || (size_t(1) << n) != len)
                throw PLError("state vector length must be a power of two");
            auto *s = new S(n, st.device_id);
            chk(b2sv_h2d(s->h, a.data(), len));
            return s;
        }))
        .def("setBasisState", [](S &s, size_t i) { chk(b2sv_set_basis_state(s.h, i)); })  // :86-91
        .def("setStateVector",                                                            // :92-108
             [](S &s, const std::vector<uint64_t> &idx,
                const py::array_t<cplx, py::array::c_style | py::array::forcecast> &v) {
                 if (idx.size() != static_cast<size_t>(v.size()))
                     throw PLError("indices and state must have the same length");
                 chk(b2sv_set_state_vector(s.h, idx.data(), reinterpret_cast<const double *>(v.data()), idx.size()));
             })
        .def("setStateOnWires",
             [](S &s, const std::vector<int64_t> &w,
                const py::array_t<cplx, py::array::c_style | py::array::forcecast> &v) {
                 if (static_cast<size_t>(v.size()) != (size_t(1) << w.size()))
                     throw PLError("state must have 2**len(wires) amplitudes");
                 chk(b2sv_set_state_on_wires(s.h, w.data(), static_cast<int>(w.size()),
                                             reinterpret_cast<const double *>(v.data())));
             })
        .def("apply",                                                                     // :233-237
             [](S &s, const std::vector<std::string> &names, const std::vector<std::vector<int64_t>> &wires,
                const std::vector<bool> &inv, const std::vector<std::vector<double>> &params) {
                 if (names.size() != wires.size())
                     throw PLError("Incompatible number of ops and wires");
                 if (names.size() != inv.size())
                     throw PLError("Incompatible number of ops and adjoints");
                 auto o = make_ops<DT>(names, params, wires, inv, {});
                 py::gil_scoped_release nogil;
                 chk(b2sv_apply_ops(s.h, o->h, 0));
             })
        .def("apply",                                                                     // :238-242
             [](S &s, const std::vector<std::string> &names, const std::vector<std::vector<int64_t>> &wires,
                const std::vector<bool> &inv) {
                 if (names.size() != wires.size())
                     throw PLError("Incompatible number of ops and wires");
                 if (names.size() != inv.size())
                     throw PLError("Incompatible number of ops and adjoints");
                 auto o = make_ops<DT>(names, std::vector<std::vector<double>>(names.size()), wires, inv, {});
                 py::gil_scoped_release nogil;
                 chk(b2sv_apply_ops(s.h, o->h, 0));
             })
        .def("apply",                                                                     // :243-261
             [](S &s, const std::string &name, const std::vector<int64_t> &w, bool inv,
                const std::vector<std::vector<double>> &, const py::array_t<cplx, py::array::c_style | py::array::forcecast> &mat) {
                 if (mat.size() == 0 || is_gate(name)) {
                     chk(b2sv_apply(s.h, name.c_str(), w.data(), static_cast<int>(w.size()), inv, nullptr, 0));
                     return;
                 }
                 if (static_cast<size_t>(mat.size()) != (size_t(1) << (2 * w.size())))
                     throw PLError("matrix size does not match the number of wires");
                 chk(b2sv_apply_matrix(s.h, w.data(), static_cast<int>(w.size()), inv,
                                       reinterpret_cast<const double *>(mat.data())));
             })
        .def("apply_ops", [](S &s, const Ops &o, bool adjoint) {
                 py::gil_scoped_release nogil;
                 chk(b2sv_apply_ops(s.h, o.h, adjoint));
             }, py::arg("ops"), py::arg("adjoint") = false)
        .def("applyGenerator",                                                            // :262-265
             [](S &s, const std::string &name, const std::vector<int64_t> &w, bool adj, const std::vector<double> &) {
                 double sc = 0;
                 chk(b2sv_apply_generator(s.h, name.c_str(), w.data(), static_cast<int>(w.size()), adj, &sc));
                 return sc;
             }, py::arg("name"), py::arg("wires"), py::arg("adjoint") = false, py::arg("params") = std::vector<double>{})
        .def("ExpectationValue",                                                          // :434-452
             [](S &s, const std::string &name, const std::vector<int64_t> &w, const std::vector<double> &,
                const py::array_t<cplx, py::array::c_style | py::array::forcecast> &mat) {
                 double out = 0;
                 if (name == "Identity" || name == "PauliX" || name == "PauliY" || name == "PauliZ" || name == "Hadamard") {
                     std::vector<int64_t> ww(w);
                     if (mat.size())
                         std::reverse(ww.begin(), ww.end()); // MeasuresKokkos.hpp:84-95
                     chk(b2sv_expval_named(s.h, name.c_str(), ww.data(), static_cast<int>(ww.size()), &out));
                 } else {
                     if (static_cast<size_t>(mat.size()) != (size_t(1) << (2 * w.size())))
                         throw PLError("matrix size does not match the number of wires");
                     chk(b2sv_expval_matrix(s.h, w.data(), static_cast<int>(w.size()),
                                            reinterpret_cast<const double *>(mat.data()), &out));
                 }
                 return out;
             })
        .def("ExpectationValue",                                                          // :453-476
             [](S &s, const std::vector<std::string> &, const std::vector<int64_t> &w, const std::vector<double> &,
                const py::array_t<cplx, py::array::c_style | py::array::forcecast> &mat) {
                 if (static_cast<size_t>(mat.size()) != (size_t(1) << (2 * w.size())))
                     throw PLError("matrix size does not match the number of wires");
                 double out = 0;
                 chk(b2sv_expval_matrix(s.h, w.data(), static_cast<int>(w.size()),
                                        reinterpret_cast<const double *>(mat.data()), &out));
                 return out;
             })
        .def("ExpectationValue",                                                          // :477-492
             [](S &s, const std::vector<int64_t> &w, const py::array_t<cplx, py::array::c_style | py::array::forcecast> &mat) {
                 if (static_cast<size_t>(mat.size()) != (size_t(1) << (2 * w.size())))
                     throw PLError("matrix size does not match the number of wires");
                 double out = 0;
                 chk(b2sv_expval_matrix(s.h, w.data(), static_cast<int>(w.size()),
                                        reinterpret_cast<const double *>(mat.data()), &out));
                 return out;
             })
        .def("ExpectationValue",                                                          // :493-516
             [](S &s, const py::array_t<cplx, py::array::c_style | py::array::forcecast> &data,
                const py::array_t<uint64_t, py::array::c_style | py::array::forcecast> &ind,
                const py::array_t<uint64_t, py::array::c_style | py::array::forcecast> &ptr) {
                 double out = 0;
                 chk(b2sv_expval_csr(s.h, reinterpret_cast<const double *>(data.data()), ind.data(), ptr.data(),
                                     static_cast<size_t>(data.size()), static_cast<size_t>(ptr.size()) - 1, &out));
                 return out;
             })
        .def("expval", [](S &s, const Obs &o) { double out = 0; chk(b2sv_expval_obs(s.h, o.h, &out)); return out; })
        .def("var", [](S &s, const Obs &o) { double out = 0; chk(b2sv_var_obs(s.h, o.h, &out)); return out; })
        .def("expval_z_all", [](S &s) {
                 py::array_t<double> out(s.nq());
                 chk(b2sv_expval_z_all(s.h, out.mutable_data(), s.nq()));
                 return out;
             })
        .def("probs", [](S &s, const std::vector<int64_t> &w) {                           // :517-533
                 const size_t m = w.empty() ? s.nq() : w.size();
                 py::array_t<double> out(size_t(1) << m);
                 chk(b2sv_probs(s.h, w.data(), static_cast<int>(w.size()), out.mutable_data()));
                 return py::array_t<np_r>(out);
             })
        .def("GenerateSamples", [](S &s, size_t num_wires, size_t shots) {                // :534-553
                 py::array_t<uint64_t> out({shots, static_cast<size_t>(s.nq())});
                 chk(b2sv_generate_samples(s.h, shots, 5374857ull, out.mutable_data()));
                 return out.reshape({shots, num_wires});
             })
        .def("DeviceToHost", [](S &s, py::array_t<np_c, py::array::c_style> &a) {         // :554-563
                 if (a.size())
                     chk(b2sv_d2h(s.h, a.mutable_data(), static_cast<size_t>(a.size())));
             })
        .def("HostToDevice", [](S &s, const arr_c &a) {                                   // :564-582
                 if (a.size())
                     chk(b2sv_h2d(s.h, a.data(), static_cast<size_t>(a.size())));
             })
        .def("numQubits", [](S &s) { return s.nq(); })
        .def("dataLength", [](S &s) { uint64_t n = 0; chk(b2sv_data_length(s.h, &n)); return n; })
        .def("resetKokkos", [](S &s) { chk(b2sv_reset(s.h)); })
        .def("set_fusion", [](S &s, bool f) { chk(b2sv_set_fusion(s.h, f)); })
        .def("reset_stats", [](S &s) { chk(b2sv_reset_stats(s.h)); })
        .def("stats", [](S &s) {
                 uint64_t a = 0, b = 0;
                 chk(b2sv_get_stats(s.h, &a, &b));
                 py::dict d;
                 d["sweeps"] = a;
                 d["launches"] = b;
                 return d;
             })
        .def("amplitudes", [](S &s, const std::vector<uint64_t> &idx) {
                 py::array_t<cplx> out(idx.size());
                 chk(b2sv_get_amplitudes(s.h, idx.data(), idx.size(), reinterpret_cast<double *>(out.mutable_data())));
                 return out;
             })
        .def("sync", [](S &s) { chk(b2sv_sync(s.h)); });
    for (const char *g : kGates) {                                                        // :109-232,266-433
        const std::string name(g);
        cls.def(g, [name](S &s, const std::vector<int64_t> &w, bool adj, const std::vector<double> &p) {
                    chk(b2sv_apply(s.h, name.c_str(), w.data(), static_cast<int>(w.size()), adj, p.data(),
                                   static_cast<int>(p.size())));
                }, py::arg("wires"), py::arg("adjoint") = false, py::arg("params") = std::vector<double>{});
    }

    // ---- observables (:591-736)
    auto obs_base = py::class_<Obs, ObsP>(m, ("ObservableKokkos_" + sfx).c_str());
    obs_base.def("__repr__", &Obs::name).def("get_wires", &Obs::wires)
        .def("__eq__", [](const Obs &a, const Obs &b) { return a.name() == b.name() && a.wires() == b.wires(); })
        .def("apply_in_place", [](const Obs &o, S &s) { chk(b2sv_obs_apply(o.h, s.h)); });
    m.def(("NamedObsKokkos_" + sfx).c_str(), [](const std::string &name, const std::vector<int64_t> &w) {
        auto o = std::make_shared<Obs>();
        chk(b2sv_obs_named(name.c_str(), w.data(), static_cast<int>(w.size()), &o->h));
        return o;
    });
    m.def(("HermitianObsKokkos_" + sfx).c_str(),
          [](const py::array_t<cplx, py::array::c_style | py::array::forcecast> &mat, const std::vector<int64_t> &w) {
              if (static_cast<size_t>(mat.size()) != (size_t(1) << (2 * w.size())))
                  throw PLError("Hermitian matrix size does not match the number of wires");
              auto o = std::make_shared<Obs>();
              chk(b2sv_obs_hermitian(reinterpret_cast<const double *>(mat.data()), w.data(), static_cast<int>(w.size()), &o->h));
              return o;
          });
    m.def(("TensorProdObsKokkos_" + sfx).c_str(), [](const std::vector<ObsP> &obs) {
        auto o = std::make_shared<Obs>();
        std::vector<b2sv_obs *> hs;
        for (const auto &c : obs)
            hs.push_back(c->h);
        chk(b2sv_obs_tensor(hs.data(), static_cast<int>(hs.size()), &o->h));
        o->children = obs;
        return o;
    });
    m.def(("HamiltonianKokkos_" + sfx).c_str(),
          [](const py::array_t<double, py::array::c_style | py::array::forcecast> &coeffs, const std::vector<ObsP> &obs) {
              if (static_cast<size_t>(coeffs.size()) != obs.size())
                  throw PLError("Assertion failed: coeffs_.size() == obs_.size()");
              auto o = std::make_shared<Obs>();
              std::vector<b2sv_obs *> hs;
              for (const auto &c : obs)
                  hs.push_back(c->h);
              chk(b2sv_obs_hamiltonian(coeffs.data(), hs.data(), static_cast<int>(hs.size()), &o->h));
              o->children = obs;
              return o;
          });
    m.def(("SparseHamiltonianKokkos_" + sfx).c_str(),
          [](const py::array_t<cplx, py::array::c_style | py::array::forcecast> &data,
             const py::array_t<uint64_t, py::array::c_style | py::array::forcecast> &ind,
             const py::array_t<uint64_t, py::array::c_style | py::array::forcecast> &ptr, const std::vector<int64_t> &w) {
              if (data.size() != ind.size())
                  throw PLError("Assertion failed: data_.size() == indices_.size()");
              auto o = std::make_shared<Obs>();
              chk(b2sv_obs_sparse(reinterpret_cast<const double *>(data.data()), ind.data(), ptr.data(),
                                  static_cast<size_t>(data.size()), static_cast<size_t>(ptr.size()) - 1, w.data(),
                                  static_cast<int>(w.size()), &o->h));
              return o;
          });

    // ---- op lists and the adjoint Jacobian (:741-821)
    py::class_<Ops, std::shared_ptr<Ops>>(m, ("OpsStructKokkos_" + sfx).c_str())
        .def(py::init([](const std::vector<std::string> &names, const std::vector<std::vector<double>> &params,
                         const std::vector<std::vector<int64_t>> &wires, const std::vector<bool> &inv) {
            return make_ops<DT>(names, params, wires, inv, {});
        }))
        .def("__len__", [](const Ops &o) { return o.names.size(); })
        .def("__repr__", [](const Ops &o) {                                               // :749-762
            std::ostringstream os;
            os << "Operations: [";
            for (size_t i = 0; i < o.names.size(); i++) {
                os << "{'name': " << o.names[i] << ", 'params': [";
                for (size_t k = 0; k < o.params[i].size(); k++)
                    os << (k ? ", " : "") << o.params[i][k];
                os << "], 'inv': " << (o.inverses[i] ? 1 : 0) << "}" << (i + 1 < o.names.size() ? "," : "");
            }
            os << "]";
            return os.str();
        });
    py::class_<Adj>(m, ("AdjointJacobianKokkos_" + sfx).c_str())
        .def(py::init<>())
        .def("create_ops_list",                                                           // :772-805
             [](Adj &, const std::vector<std::string> &names, const std::vector<py::array_t<double, py::array::c_style | py::array::forcecast>> &params,
                const std::vector<std::vector<int64_t>> &wires, const std::vector<bool> &inv,
                const std::vector<py::array_t<cplx, py::array::c_style | py::array::forcecast>> &mats) {
                 std::vector<std::vector<double>> p(params.size());
                 for (size_t i = 0; i < params.size(); i++)
                     p[i].assign(params[i].data(), params[i].data() + params[i].size());
                 std::vector<std::vector<cplx>> mm(mats.size());
                 for (size_t i = 0; i < mats.size(); i++)
                     mm[i].assign(mats[i].data(), mats[i].data() + mats[i].size());
                 return make_ops<DT>(names, p, wires, inv, mm);
             })
        .def("adjoint_jacobian",                                                          // :808-821
             [](Adj &, const S &sv, const std::vector<ObsP> &obs, const Ops &ops, const std::vector<uint64_t> &tp) {
                 std::vector<b2sv_obs *> hs;
                 for (const auto &o : obs)
                     hs.push_back(o->h);
                 py::array_t<double> jac({obs.size(), tp.size()});
                 {
                     py::gil_scoped_release nogil;
                     chk(b2sv_adjoint_jacobian(sv.h, hs.data(), static_cast<int>(hs.size()), ops.h, tp.data(),
                                               static_cast<int>(tp.size()), jac.mutable_data()));
                 }
                 return py::array_t<np_r>(jac);
             })
        .def("vjp", [](Adj &, const S &sv, const std::vector<ObsP> &obs, const Ops &ops, const std::vector<uint64_t> &tp,
                       const py::array &dy_in) {
                 if (dy_in.dtype().kind() == 'c') // lightning_kokkos.py:709-712
                     throw py::value_error("The vjp method only works with a real-valued dy when the tape is returning an expectation value");
                 const auto dy_arr = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(dy_in);
                 const std::vector<double> dy(dy_arr.data(), dy_arr.data() + dy_arr.size());
                 if (dy.size() != obs.size())
                     throw py::value_error("Number of observables in the tape must be the same as the length of dy in the vjp method");
                 std::vector<b2sv_obs *> hs;
                 for (const auto &o : obs)
                     hs.push_back(o->h);
                 py::array_t<double> out(tp.size());
                 chk(b2sv_adjoint_vjp(sv.h, hs.data(), static_cast<int>(hs.size()), dy.data(), ops.h, tp.data(),
                                      static_cast<int>(tp.size()), out.mutable_data()));
                 return py::array_t<np_r>(out);
             });
}
} // namespace

PYBIND11_MODULE(lightning_kokkos_qubit_ops_pyb, m) {
    m.doc() = "b2sv: the lightning_kokkos_qubit_ops binding surface over the B200-native engine";
    py::register_exception<PLError>(m, "PLException");                                    // :837
    m.def("kokkos_start", [] {});                                                         // :842-852
    m.def("kokkos_end", [] {});
    m.def("kokkos_config_info", [] {
        std::string buf(4096, '\0');
        chk(b2sv_backend_info(buf.data(), buf.size()));
        py::dict d;
        d["Backend"] = py::dict(py::arg("b2sv") = std::string(buf.c_str()));
        d["Version"] = std::string(b2sv_version());
        return d;
    });
    m.def("print_configuration", [] {
        std::string buf(4096, '\0');
        chk(b2sv_backend_info(buf.data(), buf.size()));
        py::print(buf.c_str());
    });
    py::class_<InitSettings>(m, "InitializationSettings")                                 // :854-967
        .def(py::init<>())
#define B2_FIELD(T, name, bitno)                                                                   \
    .def("get_" #name, [](const InitSettings &s) { return s.name; })                                \
        .def("set_" #name, [](InitSettings &s, T v) -> InitSettings & { s.name = v; s.has |= 1u << bitno; return s; }) \
        .def("has_" #name, [](const InitSettings &s) { return (s.has >> bitno & 1u) != 0; })
            B2_FIELD(int, num_threads, 0) B2_FIELD(int, device_id, 1)
                B2_FIELD(std::string, map_device_id_by, 2) B2_FIELD(bool, disable_warnings, 3)
                    B2_FIELD(bool, print_configuration, 4) B2_FIELD(bool, tune_internals, 5)
                        B2_FIELD(std::string, tools_libs, 6) B2_FIELD(bool, tools_help, 7)
                            B2_FIELD(std::string, tools_args, 8)
#undef B2_FIELD
        .def("__repr__", [](const InitSettings &s) {
            std::ostringstream os;
            os << "InitializationSettings:\nnum_threads = " << s.num_threads << "\ndevice_id = " << s.device_id;
            return os.str();
        });
    bind_precision<B2SV_C64, std::complex<float>, float>(m, "C64");
    bind_precision<B2SV_C128, std::complex<double>, double>(m, "C128");
}
