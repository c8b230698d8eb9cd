// b2sv: adjoint-method Jacobian (arXiv:2009.02823), reverse sweep over the op list.
// Reference: algorithms/AdjointDiffKokkos.hpp:404-478 (loop), :197-206 (updateJacobian).
#include "adjoint.hpp"

namespace b2sv {

void adjoint_jacobian(const State &sv, const std::vector<ObsPtr> &obs, const OpsData &ops,
                      const std::vector<uint64_t> &tp, double *jac) {
    B2_ABORT_IF(tp.empty(), "No trainable parameters provided."); // ADJ.hpp:410-411
    const size_t n_obs = obs.size(), tp_size = tp.size();
    for (size_t i = 0; i < n_obs * tp_size; i++)
        jac[i] = 0.0;

    // lambda = psi ; H_lambda[o] = O_o psi                      (ADJ.hpp:427-438)
    auto lambda = sv.clone();
    std::vector<std::unique_ptr<State>> H;
    for (size_t o = 0; o < n_obs; o++) {
        H.push_back(sv.clone());
        obs[o]->apply_in_place(*H[o]);
    }
    auto mu = sv.clone();

    long trainable_number = static_cast<long>(tp_size) - 1;
    long current_param_idx = static_cast<long>(ops.num_par_ops) - 1;
    auto tp_it = tp.rbegin();
    const auto tp_rend = tp.rend();

    for (long op_idx = static_cast<long>(ops.ops.size()) - 1; op_idx >= 0; op_idx--) {
        const GateOp &op = ops.ops[op_idx];
        B2_ABORT_IF(op.params.size() > 1, // ADJ.hpp:444-446
                    "The operation is not supported using the adjoint differentiation method");
        if (op.name == "StatePrep" || op.name == "BasisState")
            continue;
        if (tp_it == tp_rend)
            break;
        const bool has_params = !op.params.empty();
        if (has_params) {
            if (current_param_idx == static_cast<long>(*tp_it)) {
                // mu = G lambda (lambda still holds U_1..U_k psi), jac = -2 s Im<H_lambda|mu>
                mu->copy_from(*lambda);
                const double scale =
                    mu->apply_generator(op.name, op.wires) * (op.inverse ? -1.0 : 1.0);
                for (size_t o = 0; o < n_obs; o++) {
                    double im;
                    H[o]->inner_product(*mu, nullptr, &im);
                    jac[o * tp_size + trainable_number] = -2.0 * scale * im;
                }
                trainable_number--;
                ++tp_it;
            }
            current_param_idx--;
        }
        // lambda <- U^dagger lambda ; H_lambda[o] <- U^dagger H_lambda[o]
        GateOp adj = op;
        adj.inverse = !op.inverse;
        lambda->apply_gate(adj);
        for (size_t o = 0; o < n_obs; o++)
            H[o]->apply_gate(adj);
    }
    lambda->sync();
    for (auto &h : H)
        h->sync();
}

} // namespace b2sv
