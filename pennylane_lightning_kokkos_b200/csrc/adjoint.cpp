// b2sv: adjoint-method Jacobian (arXiv:2009.02823), reverse sweep over the op list.
// Reference: algorithms/AdjointDiffKokkos.hpp:404-478 (loop), :197-206 (updateJacobian).
//
// Same recurrence as the reference (lambda, H_lambda[o], one Jacobian column per trainable op), but
//   * every Jacobian entry is reduced INTO a device-resident array and read back once at the end
//     (the reference blocks on one scalar per (observable, parameter), ADJ.hpp:202-205);
//   * for Pauli-word generators (RX, RY, RZ, IsingXX/YY/ZZ, MultiRZ)  Im<H_lambda| G |lambda>  is
//     taken in one read pass over the two vectors -- no mu = lambda copy, no generator sweep
//     (reference: copy 2S + generator 2S + inner product 2S, ADJ.hpp:454-470);
//   * the U^dagger updates of lambda and H_lambda between two trainable ops go through the fusing
//     tile executor as one batch.
// All states of one call live on one stream, so nothing synchronises until the final read-back.
#include "adjoint.hpp"

namespace b2sv {

void adjoint_jacobian(const State &sv, const std::vector<ObsPtr> &obs, const OpsData &ops,
                      const std::vector<uint64_t> &tp, double *jac) {
    B2_ABORT_IF(tp.empty(), "No trainable parameters provided."); // ADJ.hpp:410-411
    const size_t n_obs = obs.size(), tp_size = tp.size();
    for (size_t i = 0; i < n_obs * tp_size; i++)
        jac[i] = 0.0;
    for (const GateOp &op : ops.ops) // ADJ.hpp:444-446 (checked up front: nothing is launched)
        B2_ABORT_IF(op.params.size() > 1,
                    "The operation is not supported using the adjoint differentiation method");
    if (n_obs == 0)
        return;
    CUDA_CHECK(cudaSetDevice(sv.device()));

    // lambda = psi ; H_lambda[o] = O_o psi                      (ADJ.hpp:427-438)
    auto lambda = sv.clone_on_stream();
    std::vector<std::unique_ptr<State>> H;
    for (size_t o = 0; o < n_obs; o++) {
        H.push_back(lambda->clone_on_stream());
        obs[o]->apply_in_place(*H[o]);
    }
    std::unique_ptr<State> mu; // only for generators that are not Pauli words
    cudaStream_t st = lambda->stream();
    double *d_jac = nullptr;
    CUDA_CHECK(cudaMalloc(&d_jac, sizeof(double) * n_obs * tp_size));
    CUDA_CHECK(cudaMemsetAsync(d_jac, 0, sizeof(double) * n_obs * tp_size, st));

    long trainable_number = static_cast<long>(tp_size) - 1;
    long current_param_idx = static_cast<long>(ops.num_par_ops) - 1;
    auto tp_it = tp.rbegin();
    const auto tp_rend = tp.rend();
    static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};

    std::vector<GateOp> batch; // U^dagger ops not yet applied to lambda / H_lambda, in order
    auto flush = [&]() {
        if (batch.empty())
            return;
        lambda->apply_ops(batch, false);
        for (auto &h : H)
            h->apply_ops(batch, false);
        batch.clear();
    };

    for (long op_idx = static_cast<long>(ops.ops.size()) - 1; op_idx >= 0; op_idx--) {
        const GateOp &op = ops.ops[op_idx];
        if (op.name == "StatePrep" || op.name == "BasisState") // ADJ.hpp:447-450
            continue;
        if (tp_it == tp_rend) // ADJ.hpp:451-453
            break;
        if (!op.params.empty()) {
            if (current_param_idx == static_cast<long>(*tp_it)) {
                flush(); // lambda and H_lambda now hold the state right after this op
                const std::vector<int> bits = wires_to_bits(op.wires, sv.num_qubits());
                uint64_t x, z;
                int ny;
                double scale;
                const double sign = op.inverse ? -1.0 : 1.0;
                if (generator_pauli(op.name, bits, &x, &z, &ny, &scale)) {
                    for (size_t o = 0; o < n_obs; o++)
                        lambda->pauli_dot_im_to(*H[o], x, z, ipow[ny & 3], -2.0 * scale * sign,
                                                d_jac + o * tp_size + trainable_number);
                } else { // mu = G lambda, jac = -2 s Im<H_lambda|mu>     (ADJ.hpp:454-470)
                    if (!mu)
                        mu = lambda->clone_on_stream();
                    else
                        mu->copy_from(*lambda);
                    scale = mu->apply_generator(op.name, op.wires);
                    for (size_t o = 0; o < n_obs; o++)
                        mu->dot_im_to(*H[o], -2.0 * scale * sign,
                                      d_jac + o * tp_size + trainable_number);
                }
                trainable_number--;
                ++tp_it;
            }
            current_param_idx--;
        }
        // lambda <- U^dagger lambda ; H_lambda[o] <- U^dagger H_lambda[o]   (ADJ.hpp:455,476)
        GateOp adj = op;
        adj.inverse = !op.inverse;
        batch.push_back(std::move(adj));
    }
    // ops before the first trainable one never influence the Jacobian: the batch is dropped
    lambda->allreduce_device(d_jac, static_cast<int>(n_obs * tp_size));
    CUDA_CHECK(cudaMemcpyAsync(jac, d_jac, sizeof(double) * n_obs * tp_size, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    CUDA_CHECK(cudaFree(d_jac));
}

} // namespace b2sv
