// b2sv: observables (see obs.hpp). Reference: simulator/ObservablesKokkos.hpp.
//
// The reference applies a Hamiltonian term by term: copy the state, apply the term, axpy into a
// zeroed buffer (OBS.hpp:360-373: ~(5+2p) S of traffic per term and two temporaries). Here a
// Hamiltonian whose terms are Pauli words is applied by ONE kernel that gathers psi[i ^ x_t] for
// every term and writes each output amplitude once (kernels.cu k_pauli_sum_apply).
#include "obs.hpp"

#include <algorithm>
#include <set>
#include <sstream>

namespace b2sv {
namespace {

std::string wires_str(const std::vector<int64_t> &w) { // Util operator<< for vectors
    std::ostringstream os;
    os << '[';
    for (size_t i = 0; i < w.size(); i++) {
        if (i)
            os << ", ";
        os << w[i];
    }
    os << ']';
    return os.str();
}

class NamedObs final : public Obs {
    std::string name_;
    std::vector<int64_t> wires_;

  public:
    NamedObs(std::string n, std::vector<int64_t> w) : name_(std::move(n)), wires_(std::move(w)) {}
    void apply_in_place(State &sv) const override { // OBS.hpp:116-118
        GateOp op;
        op.name = name_;
        op.wires = wires_;
        B2_ABORT_IF(name_ != "Identity" && !is_named_gate(name_),
                    "observable '" + name_ + "' is not a named gate");
        sv.apply_gate(op);
    }
    std::string name() const override { return name_ + wires_str(wires_); }
    std::vector<int64_t> wires() const override { return wires_; }
    bool pauli_terms(int n, double coef,
                     std::vector<std::pair<double, PauliWord>> &out) const override {
        PauliWord w;
        if (name_ != "Identity") {
            if (wires_.size() != 1 || wires_[0] < 0 || wires_[0] >= n)
                return false;
            const uint64_t b = bit(n - 1 - static_cast<int>(wires_[0]));
            if (name_ == "PauliX") {
                w.x = b;
            } else if (name_ == "PauliY") {
                w.x = b;
                w.z = b;
                w.ny = 1;
            } else if (name_ == "PauliZ") {
                w.z = b;
            } else {
                return false;
            }
        }
        out.emplace_back(coef, w);
        return true;
    }
};

class HermitianObs final : public Obs {
    std::vector<cplx> matrix_;
    std::vector<int64_t> wires_;

  public:
    HermitianObs(std::vector<cplx> m, std::vector<int64_t> w)
        : matrix_(std::move(m)), wires_(std::move(w)) {}
    void apply_in_place(State &sv) const override { // OBS.hpp:173-179: the matrix path
        GateOp op;
        op.name = "Hermitian";
        op.wires = wires_;
        op.matrix = matrix_;
        sv.apply_gate(op);
    }
    std::string name() const override {
        // reference: "Hermitian" + MatrixHasher(matrix) (OBS.hpp:160-167); same idea, own hash
        uint64_t h = 1469598103934665603ull;
        for (const cplx &c : matrix_) {
            const double v[2] = {c.real(), c.imag()};
            const unsigned char *p = reinterpret_cast<const unsigned char *>(v);
            for (size_t i = 0; i < sizeof(v); i++)
                h = (h ^ p[i]) * 1099511628211ull;
        }
        std::ostringstream os;
        os << "Hermitian" << h;
        return os.str();
    }
    std::vector<int64_t> wires() const override { return wires_; }
};

class TensorProdObs final : public Obs {
    std::vector<ObsPtr> obs_;
    std::vector<int64_t> all_wires_;

  public:
    explicit TensorProdObs(std::vector<ObsPtr> obs) : obs_(std::move(obs)) {
        std::set<int64_t> seen; // OBS.hpp:214-228
        for (const auto &ob : obs_)
            for (int64_t w : ob->wires()) {
                B2_ABORT_IF(seen.count(w), "All wires in observables must be disjoint.");
                seen.insert(w);
            }
        all_wires_.assign(seen.begin(), seen.end());
    }
    void apply_in_place(State &sv) const override {
        for (const auto &ob : obs_)
            ob->apply_in_place(sv);
    }
    std::string name() const override {
        std::string s;
        for (size_t i = 0; i < obs_.size(); i++) {
            s += obs_[i]->name();
            if (i + 1 != obs_.size())
                s += " @ ";
        }
        return s;
    }
    std::vector<int64_t> wires() const override { return all_wires_; }
    bool pauli_terms(int n, double coef,
                     std::vector<std::pair<double, PauliWord>> &out) const override {
        PauliWord acc;
        for (const auto &ob : obs_) {
            std::vector<std::pair<double, PauliWord>> one;
            if (!ob->pauli_terms(n, 1.0, one) || one.size() != 1)
                return false;
            coef *= one[0].first;     // a factor may be a one-term Hamiltonian with its own weight
            acc.x |= one[0].second.x; // wires are disjoint
            acc.z |= one[0].second.z;
            acc.ny += one[0].second.ny;
        }
        out.emplace_back(coef, acc);
        return true;
    }
};

class HamiltonianObs final : public Obs {
    std::vector<double> coeffs_;
    std::vector<ObsPtr> obs_;

  public:
    HamiltonianObs(std::vector<double> c, std::vector<ObsPtr> o)
        : coeffs_(std::move(c)), obs_(std::move(o)) {
        B2_ASSERT(coeffs_.size() == obs_.size());
    }
    void apply_in_place(State &sv) const override;
    std::string name() const override { // OBS.hpp:387-400
        std::ostringstream ss;
        ss << "Hamiltonian: { 'coeffs' : [";
        for (size_t i = 0; i < coeffs_.size(); i++) {
            if (i)
                ss << ", ";
            ss << coeffs_[i];
        }
        ss << "], 'observables' : [";
        for (size_t t = 0; t < obs_.size(); t++) {
            ss << obs_[t]->name();
            if (t + 1 != obs_.size())
                ss << ", ";
        }
        ss << "]}";
        return ss.str();
    }
    std::vector<int64_t> wires() const override {
        std::set<int64_t> s;
        for (const auto &ob : obs_)
            for (int64_t w : ob->wires())
                s.insert(w);
        return std::vector<int64_t>(s.begin(), s.end());
    }
    bool pauli_terms(int n, double coef,
                     std::vector<std::pair<double, PauliWord>> &out) const override {
        std::vector<std::pair<double, PauliWord>> tmp;
        for (size_t i = 0; i < obs_.size(); i++)
            if (!obs_[i]->pauli_terms(n, coef * coeffs_[i], tmp))
                return false;
        out.insert(out.end(), tmp.begin(), tmp.end());
        return true;
    }
};

class SparseHamiltonianObs final : public Obs {
    std::vector<cplx> data_;
    std::vector<uint64_t> indices_, indptr_;
    std::vector<int64_t> wires_;
    mutable std::shared_ptr<CsrDevice> dev_; // uploaded once, reused (reference re-uploads)

  public:
    SparseHamiltonianObs(std::vector<cplx> d, std::vector<uint64_t> i, std::vector<uint64_t> p,
                         std::vector<int64_t> w)
        : data_(std::move(d)), indices_(std::move(i)), indptr_(std::move(p)), wires_(std::move(w)) {
        B2_ASSERT(data_.size() == indices_.size());
        B2_ABORT_IF(indptr_.empty(), "CSR indptr must not be empty");
    }
    const CsrDevice &device_csr(int device) const {
        if (!dev_ || dev_->device != device)
            dev_ = csr_upload(device, data_.data(), indices_.data(), indptr_.data(), data_.size(),
                              indptr_.size() - 1);
        return *dev_;
    }
    void apply_in_place(State &sv) const override { // OBS.hpp:484-494
        B2_ABORT_IF(static_cast<int>(wires_.size()) != sv.num_qubits(),
                    "SparseH wire count does not match state-vector size");
        B2_ABORT_IF(sv.world() > 1, "SparseHamiltonian::applyInPlace is not supported on sharded states (expval is)");
        const CsrDevice &m = device_csr(sv.device());
        B2_ABORT_IF(m.nrows != sv.local_length(), "CSR matrix dimension does not match the state vector");
        void *y = sv.acquire_scratch();
        if (sv.alloc_length() != sv.local_length())
            CUDA_CHECK(cudaMemsetAsync(y, 0, sv.alloc_length() * sv.amp_bytes(), sv.stream()));
        launch_csr_spmv(sv.dtype(), sv.data(), y, m.data, m.ind, m.ptr, m.nrows, m.lanes, sv.stream());
        sv.launches++;
        sv.bytes_moved += m.nnz * 20 + 2 * sv.state_bytes();
        sv.swap_buffer(y);
        sv.release_scratch(y);
    }
    double expval(const State &sv) const { return sv.expval_csr(device_csr(sv.device())); }
    std::string name() const override { // OBS.hpp:496-512
        std::ostringstream ss;
        ss << "SparseHamiltonian: {\n'data' : ";
        for (const auto &d : data_)
            ss << d;
        ss << ",\n'indices' : ";
        for (auto i : indices_)
            ss << i;
        ss << ",\n'indptr' : ";
        for (auto o : indptr_)
            ss << o;
        ss << "\n}";
        return ss.str();
    }
    std::vector<int64_t> wires() const override { return wires_; }
};

// out_buffer <- sum_t c_t P_t |in>; both raw device buffers of `sv` geometry
void pauli_sum_into(const State &sv, const void *in, void *out,
                    const std::vector<std::pair<double, PauliWord>> &terms) {
    B2_ABORT_IF(sv.world() > 1, "Pauli-sum application on sharded states is not supported yet");
    std::vector<PauliTerm> h(terms.size());
    static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
    for (size_t t = 0; t < terms.size(); t++) {
        const cplx c = terms[t].first * ipow[terms[t].second.ny & 3];
        h[t] = PauliTerm{terms[t].second.x, terms[t].second.z, c.real(), c.imag(), 0, 0};
    }
    PauliTerm *d_terms;
    CUDA_CHECK(cudaSetDevice(sv.device()));
    CUDA_CHECK(cudaMallocAsync(&d_terms, sizeof(PauliTerm) * h.size(), sv.stream()));
    CUDA_CHECK(cudaMemcpyAsync(d_terms, h.data(), sizeof(PauliTerm) * h.size(),
                               cudaMemcpyHostToDevice, sv.stream()));
    launch_pauli_sum_apply(sv.dtype(), in, out, sv.local_length(), d_terms,
                           static_cast<int>(h.size()), sv.stream());
    { // one read of the input per distinct x mask (the others hit the same lines) + one write
        std::set<uint64_t> xs;
        for (const PauliTerm &t : h)
            xs.insert(t.x);
        sv.bytes_moved += (xs.size() + 1) * sv.state_bytes();
    }
    CUDA_CHECK(cudaFreeAsync(d_terms, sv.stream()));
    CUDA_CHECK(cudaStreamSynchronize(sv.stream())); // `h` goes out of scope
}

void HamiltonianObs::apply_in_place(State &sv) const {
    std::vector<std::pair<double, PauliWord>> terms;
    if (sv.world() > 1 && pauli_terms(sv.num_qubits(), 1.0, terms)) {
        // sharded: one kernel, partner amplitudes of terms that flip a global qubit come straight from
        // the peer's shard over NVLink
        static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        std::vector<PauliTerm> pt(terms.size());
        for (size_t t = 0; t < terms.size(); t++) {
            const cplx c = terms[t].first * ipow[terms[t].second.ny & 3];
            pt[t] = PauliTerm{terms[t].second.x, terms[t].second.z, c.real(), c.imag(), 0, 0};
        }
        if (sv.pauli_sum_apply_sharded(pt))
            return;
        terms.clear();
    }
    if (sv.world() == 1 && pauli_terms(sv.num_qubits(), 1.0, terms)) {
        void *out = sv.acquire_scratch();
        if (sv.alloc_length() != sv.local_length())
            CUDA_CHECK(cudaMemsetAsync(out, 0, sv.alloc_length() * sv.amp_bytes(), sv.stream()));
        pauli_sum_into(sv, sv.data(), out, terms);
        sv.launches++;
        sv.swap_buffer(out);
        sv.release_scratch(out);
        return;
    }
    // generic path, as the reference (OBS.hpp:360-373): buffer = sum_t c_t (O_t sv)
    auto buffer = sv.clone();
    buffer->init_zeros();
    auto tmp = sv.clone();
    for (size_t t = 0; t < coeffs_.size(); t++) {
        if (t)
            tmp->copy_from(sv);
        obs_[t]->apply_in_place(*tmp);
        buffer->axpy(cplx(coeffs_[t], 0.0), *tmp);
    }
    sv.copy_from(*buffer);
    sv.sync();
}

} // namespace

ObsPtr make_named_obs(const std::string &name, const std::vector<int64_t> &wires) {
    return std::make_shared<NamedObs>(name, wires);
}
ObsPtr make_hermitian_obs(const std::vector<cplx> &matrix, const std::vector<int64_t> &wires) {
    B2_ABORT_IF(matrix.size() != (size_t(1) << (2 * wires.size())),
                "Hermitian matrix size does not match the number of wires");
    return std::make_shared<HermitianObs>(matrix, wires);
}
ObsPtr make_tensor_obs(const std::vector<ObsPtr> &obs) { return std::make_shared<TensorProdObs>(obs); }
ObsPtr make_hamiltonian_obs(const std::vector<double> &coeffs, const std::vector<ObsPtr> &obs) {
    return std::make_shared<HamiltonianObs>(coeffs, obs);
}
ObsPtr make_sparse_obs(const std::vector<cplx> &data, const std::vector<uint64_t> &indices,
                       const std::vector<uint64_t> &indptr, const std::vector<int64_t> &wires) {
    return std::make_shared<SparseHamiltonianObs>(data, indices, indptr, wires);
}

void apply_obs_to_buffer(const State &sv, const Obs &ob, State &out) {
    out.copy_from(sv);
    ob.apply_in_place(out);
}

double expval_obs(const State &sv, const Obs &ob) {
    // fast paths first; all equal Re<psi|O psi> of the reference (MK.hpp:354-360)
    if (auto *sp = dynamic_cast<const SparseHamiltonianObs *>(&ob))
        return sp->expval(sv);
    std::vector<std::pair<double, PauliWord>> terms;
    if (ob.pauli_terms(sv.num_qubits(), 1.0, terms) && terms.size() == 1) {
        static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        const PauliWord &w = terms[0].second;
        return terms[0].first * sv.expval_pauli(w.x, w.z, ipow[w.ny & 3]);
    }
    if (!terms.empty() && sv.world() == 1) { // a sum of Pauli words: one read pass per distinct x mask, no work vector
        static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        std::vector<PauliTerm> pt(terms.size());
        for (size_t t = 0; t < terms.size(); t++) {
            const cplx c = terms[t].first * ipow[terms[t].second.ny & 3];
            pt[t] = PauliTerm{terms[t].second.x, terms[t].second.z, c.real(), c.imag(), 0, 0};
        }
        return sv.expval_pauli_sum(pt);
    }
    auto tmp = sv.clone();
    ob.apply_in_place(*tmp);
    double re;
    sv.inner_product(*tmp, &re, nullptr);
    return re;
}

double var_obs(const State &sv, const Obs &ob) { // MK.hpp:368-381
    auto tmp = sv.clone();
    ob.apply_in_place(*tmp);
    double sq, mean;
    tmp->inner_product(*tmp, &sq, nullptr);
    sv.inner_product(*tmp, &mean, nullptr);
    return sq - mean * mean;
}

} // namespace b2sv
