// b2sv: observable object model with applyInPlace semantics.
// Counterpart of reference simulator/ObservablesKokkos.hpp:19-517.
#pragma once
#include "state.hpp"

namespace b2sv {

struct PauliWord {
    uint64_t x = 0, z = 0; // X or Y on bit -> x ; Z or Y on bit -> z
    int ny = 0;            // number of Y factors: P|j> = i^ny (-1)^popc(j&z) |j^x>
};

class Obs {
  public:
    virtual ~Obs() = default;
    virtual void apply_in_place(State &sv) const = 0;           // OBS.hpp:48
    virtual std::string name() const = 0;                        // getObsName
    virtual std::vector<int64_t> wires() const = 0;              // getWires
    // Fast path: the observable as a real-weighted sum of Pauli words (Identity allowed).
    virtual bool pauli_terms(int num_qubits, double coef,
                             std::vector<std::pair<double, PauliWord>> &out) const {
        (void)num_qubits; (void)coef; (void)out;
        return false;
    }
};
using ObsPtr = std::shared_ptr<const Obs>;

ObsPtr make_named_obs(const std::string &name, const std::vector<int64_t> &wires);
ObsPtr make_hermitian_obs(const std::vector<cplx> &matrix, const std::vector<int64_t> &wires);
ObsPtr make_tensor_obs(const std::vector<ObsPtr> &obs);
ObsPtr make_hamiltonian_obs(const std::vector<double> &coeffs, const std::vector<ObsPtr> &obs);
ObsPtr make_sparse_obs(const std::vector<cplx> &data, const std::vector<uint64_t> &indices,
                       const std::vector<uint64_t> &indptr, const std::vector<int64_t> &wires);

// MeasuresKokkos::expval(ob) / var(ob) (reference MeasuresKokkos.hpp:354-381)
double expval_obs(const State &sv, const Obs &ob);
double var_obs(const State &sv, const Obs &ob);
// out <- O |sv> without touching sv; `out` is a raw buffer of sv.alloc_length() amplitudes
void apply_obs_to_buffer(const State &sv, const Obs &ob, State &out);

} // namespace b2sv
