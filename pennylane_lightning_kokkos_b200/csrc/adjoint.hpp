// b2sv: adjoint-method Jacobian. Counterpart of reference algorithms/AdjointDiffKokkos.hpp.
#pragma once
#include "obs.hpp"

namespace b2sv {

struct OpsData { // reference AdjointDiffKokkos.hpp:17-173
    std::vector<GateOp> ops;
    size_t num_par_ops = 0;
};

// jac: row-major n_obs x tp.size(), overwritten. Reference: AdjointDiffKokkos.hpp:404-478.
void adjoint_jacobian(const State &sv, const std::vector<ObsPtr> &obs, const OpsData &ops,
                      const std::vector<uint64_t> &trainable, double *jac);

} // namespace b2sv
