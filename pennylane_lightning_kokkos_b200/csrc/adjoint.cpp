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
//     tile executor as one batch;
//   * a RUN of consecutive single-qubit gates (any wires; e.g. the RX/RY/RZ block of one ansatz
//     layer) is differentiated without undoing its gates one by one: with W_k the product of the
//     run's gates after gate k,  <H_lambda_k| G_k |lambda_k> = <H_lambda_b| W_k G_k W_k^dagger |lambda_b>,
//     and because gates on other wires cancel, W_k G_k W_k^dagger is again a 2x2 on gate k's wire
//     (computed on the host). So the whole run needs, per wire, the four transition sums
//     <H_lambda_b| |i><j|_t |lambda_b>  -- one read pass per six wires (kernels.cu k_transition_1q) --
//     and its U^dagger updates reach the states as ONE fused batch together with whatever follows.
//     Config 3 (24 qubits, 7 layers of 72 rotations): ~0.4 TB of traffic instead of ~3.3 TB.
// All states of one call live on one stream, so nothing synchronises until the final read-back.
#include "adjoint.hpp"

#include "kernels.cuh"

#include <algorithm>
#include <array>
#include <cstdlib>

namespace b2sv {

void adjoint_jacobian(const State &sv, const std::vector<ObsPtr> &obs, const OpsData &ops,
                      const std::vector<uint64_t> &tp, double *jac) {
    B2_ABORT_IF(tp.empty(), "No trainable parameters provided."); // ADJ.hpp:410-411
    const size_t n_obs = obs.size(), tp_size = tp.size();
    for (size_t i = 0; i < n_obs * tp_size; i++)
        jac[i] = 0.0;
    { // ADJ.hpp:444-453: the reference checks inside the reverse loop, so ops in front of the point
      // where the trainable parameters run out are never looked at. Replay the loop's bookkeeping
      // on the host first: exactly the ops the loop visits are validated, nothing is launched
      // when one fails.
        auto it = tp.rbegin();
        long cur = static_cast<long>(ops.num_par_ops) - 1;
        for (size_t i = ops.ops.size(); i-- > 0;) {
            const GateOp &op = ops.ops[i];
            B2_ABORT_IF(op.params.size() > 1,
                        "The operation is not supported using the adjoint differentiation method");
            if (op.name == "StatePrep" || op.name == "BasisState")
                continue;
            if (it == tp.rend())
                break;
            if (!op.params.empty()) {
                if (cur == static_cast<long>(*it))
                    ++it;
                cur--;
            }
        }
    }
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
    if (sv.sharded()) {
        // applying an observable may leave H_lambda in another qubit layout than lambda (a Hamiltonian
        // built term by term starts from a fresh buffer); the sweep needs them identical, and they
        // stay identical from here on because every state sees the same ops
        bool same = true;
        for (auto &h : H)
            same = same && lambda->same_layout(*h);
        if (!same) {
            lambda->normalize_layout();
            for (auto &h : H)
                h->normalize_layout();
        }
    }
    std::unique_ptr<State> mu; // only for generators that are not Pauli words
    cudaStream_t st = lambda->stream();
    double *d_jac = nullptr;
    CUDA_CHECK(cudaMallocAsync(&d_jac, sizeof(double) * n_obs * tp_size, st));
    CUDA_CHECK(cudaMemsetAsync(d_jac, 0, sizeof(double) * n_obs * tp_size, st));

    long trainable_number = static_cast<long>(tp_size) - 1;
    long current_param_idx = static_cast<long>(ops.num_par_ops) - 1;
    auto tp_it = tp.rbegin();
    const auto tp_rend = tp.rend();
    static const cplx ipow[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};

    std::vector<GateOp> batch; // U^dagger ops not yet applied to lambda / H_lambda, in order
    std::vector<std::pair<size_t, size_t>> run_segments; // [begin, end) of runs inside `batch`
    auto flush = [&]() {
        if (batch.empty())
            return;
        // wires are validated when the ops are lowered; an out-of-range wire must reach that check
        // (and its error message) instead of indexing the tables below
        for (const GateOp &op : batch)
            for (auto w : op.wires)
                if (w < 0 || w >= static_cast<int64_t>(sv.num_qubits()))
                    run_segments.clear();
        // The single-qubit gates of a run commute across wires, so their order inside the batch is
        // free. Put them in the order in which the ops AFTER the run first act on their wires
        // (last wire of a multi-qubit gate = its target): the tile passes that carry the gates then
        // also absorb the permutation gates that follow, instead of leaving them to extra sweeps
        // (reversed CNOT ring after a rotation block: 3 passes instead of 6).
        for (const auto &seg : run_segments) {
            std::vector<size_t> key(sv.num_qubits(), batch.size());
            for (size_t i = batch.size(); i-- > seg.second;)
                if (!batch[i].wires.empty()) {
                    if (batch[i].wires.size() == 1)
                        key[batch[i].wires[0]] = i;
                    else
                        key[batch[i].wires.back()] = i;
                }
            std::stable_sort(batch.begin() + seg.first, batch.begin() + seg.second,
                             [&](const GateOp &a, const GateOp &b) { return key[a.wires[0]] < key[b.wires[0]]; });
        }
        // Then let every later op slide forward to just behind the last op it shares a wire with
        // (ops on disjoint wires commute): the entangling gates end up interleaved with the run's
        // gates the way a forward circuit has them, which is the order the pass scheduler fuses best.
        if (!run_segments.empty() && batch.size() <= 4096) {
            std::vector<GateOp> out;
            out.reserve(batch.size());
            std::vector<long> last_on_wire(sv.num_qubits(), -1); // position in `out`
            for (size_t i = 0; i < batch.size(); i++) {
                bool in_run = false;
                for (const auto &seg : run_segments)
                    in_run = in_run || (i >= seg.first && i < seg.second);
                if (i < run_segments.front().first || in_run) { // these keep their (sorted) order
                    for (auto w : batch[i].wires)
                        last_on_wire[w] = static_cast<long>(out.size());
                    out.push_back(batch[i]);
                    continue;
                }
                long dep = static_cast<long>(run_segments.front().first) - 1;
                for (auto w : batch[i].wires)
                    dep = std::max(dep, last_on_wire[w]);
                const size_t pos = static_cast<size_t>(dep + 1);
                out.insert(out.begin() + pos, batch[i]);
                for (long &l : last_on_wire)
                    if (l >= static_cast<long>(pos))
                        l++;
                for (auto w : batch[i].wires)
                    last_on_wire[w] = static_cast<long>(pos);
            }
            batch.swap(out);
        }
        run_segments.clear();
        std::vector<State *> all = {lambda.get()};
        for (auto &h : H)
            all.push_back(h.get());
        State::apply_ops_to_all(all, batch); // lowered and scheduled once
        batch.clear();
    };

    // ---- runs of single-qubit gates (see the header comment)
    struct RunOp {
        const GateOp *op;
        int bit;          // index bit of the gate's wire
        cplx u[4];        // the gate as applied in the circuit (row-major 2x2)
        long slot;        // Jacobian column, -1 = not trainable
        cplx g[4];        // generator (trainable ops)
        double coeff;     // -2 * scale * sign
    };
    std::vector<RunOp> run; // in processing order = descending op index
    std::vector<double> jac_host(n_obs * tp_size, 0.0);
    // B2SV_ADJOINT_RUNS=0 turns the run path off (A/B measurements). Sharded states take the run path too:
    // the wires of a chunk are brought into the shard and the transition sums are all-reduced.
    const char *runs_env = getenv("B2SV_ADJOINT_RUNS");
    const bool runs_enabled = runs_env == nullptr || atoi(runs_env) != 0;
    double *d_tr_scratch = nullptr, *d_tr_out = nullptr;
    auto one_qubit_matrix = [&](const GateOp &op, int *bit, cplx u[4]) {
        if (op.wires.size() != 1 || !op.matrix.empty())
            return false;
        std::vector<Prim> prims;
        const std::vector<int> bits = wires_to_bits(op.wires, sv.num_qubits());
        if (!lower_gate(op.name, bits, op.inverse, op.params, prims) || prims.size() != 1)
            return false;
        const Prim &p = prims[0];
        if (p.cmask != 0)
            return false;
        if (p.type == Prim::C1Q) {
            for (int i = 0; i < 4; i++)
                u[i] = p.m[i];
        } else if (p.type == Prim::DIAG && __builtin_popcountll(p.pmask) == 1) {
            u[0] = p.m[0];
            u[1] = u[2] = 0.0;
            u[3] = p.m[1];
        } else {
            return false;
        }
        *bit = bits[0];
        return true;
    };
    auto generator_2x2 = [](const std::string &name, cplx g[4], double *scale) {
        const cplx I(0.0, 1.0);
        if (name == "RX") {
            g[0] = 0, g[1] = 1, g[2] = 1, g[3] = 0, *scale = -0.5;
        } else if (name == "RY") {
            g[0] = 0, g[1] = -I, g[2] = I, g[3] = 0, *scale = -0.5;
        } else if (name == "RZ") {
            g[0] = 1, g[1] = 0, g[2] = 0, g[3] = -1, *scale = -0.5;
        } else if (name == "PhaseShift") { // SV.hpp:1275-1281: projector |1><1|, factor 1
            g[0] = 0, g[1] = 0, g[2] = 0, g[3] = 1, *scale = 1.0;
        } else {
            return false;
        }
        return true;
    };
    auto mul2 = [](const cplx a[4], const cplx b[4], cplx c[4]) {
        const cplx r0 = a[0] * b[0] + a[1] * b[2], r1 = a[0] * b[1] + a[1] * b[3];
        const cplx r2 = a[2] * b[0] + a[3] * b[2], r3 = a[2] * b[1] + a[3] * b[3];
        c[0] = r0, c[1] = r1, c[2] = r2, c[3] = r3;
    };
    auto finish_run = [&]() {
        if (run.empty())
            return;
        bool any = false;
        for (const RunOp &r : run)
            any = any || r.slot >= 0;
        if (any) {
            flush(); // lambda, H_lambda = the states right after the last gate of the run
            // W G W^dagger per trainable gate; W = product of the run's later gates on the same wire
            std::vector<std::array<cplx, 4>> acc(sv.num_qubits(), {cplx(1), cplx(0), cplx(0), cplx(1)});
            struct Need {
                int bit;
                cplx m[4];
                double coeff;
                long slot;
            };
            std::vector<Need> needs;
            std::vector<int> wires_needed;
            for (const RunOp &r : run) {
                cplx *a = acc[r.bit].data();
                if (r.slot >= 0) {
                    Need nd;
                    nd.bit = r.bit;
                    nd.coeff = r.coeff;
                    nd.slot = r.slot;
                    cplx t[4];
                    const cplx ad[4] = {std::conj(a[0]), std::conj(a[2]), std::conj(a[1]), std::conj(a[3])};
                    mul2(a, r.g, t);
                    mul2(t, ad, nd.m);
                    needs.push_back(nd);
                    if (std::find(wires_needed.begin(), wires_needed.end(), r.bit) == wires_needed.end())
                        wires_needed.push_back(r.bit);
                }
                mul2(a, r.u, a);
            }
            if (!d_tr_scratch) {
                CUDA_CHECK(cudaMallocAsync(&d_tr_scratch, sizeof(double) * kReduceBlocks * kTransitionVals, st));
                CUDA_CHECK(cudaMallocAsync(&d_tr_out, sizeof(double) * kTransitionVals, st));
            }
            std::vector<double> vals(kTransitionVals);
            for (size_t c0 = 0; c0 < wires_needed.size(); c0 += kTransitionBits) {
                const int nb = static_cast<int>(std::min<size_t>(kTransitionBits, wires_needed.size() - c0));
                // sharded states: the chunk's qubits come into the shard on lambda and on every H_lambda
                // (identical layouts, hence identical exchanges), then each rank sums over its shard
                int phys[kTransitionBits];
                if (sv.sharded()) {
                    uint64_t mask = 0;
                    for (int t = 0; t < nb; t++)
                        mask |= bit(wires_needed[c0 + t]);
                    lambda->ensure_local(mask);
                    for (auto &h : H)
                        h->ensure_local(mask);
                }
                for (int t = 0; t < nb; t++)
                    phys[t] = lambda->phys_bit(wires_needed[c0 + t]);
                for (size_t o = 0; o < n_obs; o++) {
                    lambda->transition_1q_to(*H[o], phys, nb, d_tr_scratch, d_tr_out);
                    CUDA_CHECK(cudaMemcpyAsync(vals.data(), d_tr_out, sizeof(double) * kTransitionVals,
                                               cudaMemcpyDeviceToHost, st));
                    CUDA_CHECK(cudaStreamSynchronize(st));
                    const cplx D(vals[0], vals[1]);
                    for (int t = 0; t < nb; t++) {
                        const cplx Z(vals[2 + 6 * t], vals[3 + 6 * t]), X(vals[4 + 6 * t], vals[5 + 6 * t]),
                            W(vals[6 + 6 * t], vals[7 + 6 * t]);
                        // E_ij = <H_lambda| (|i><j|)_t |lambda>
                        const cplx E00 = 0.5 * (D + Z), E11 = 0.5 * (D - Z), E01 = 0.5 * (X + W),
                                   E10 = 0.5 * (X - W);
                        for (const Need &nd : needs)
                            if (nd.bit == wires_needed[c0 + t]) {
                                const cplx v = nd.m[0] * E00 + nd.m[1] * E01 + nd.m[2] * E10 + nd.m[3] * E11;
                                jac_host[o * tp_size + nd.slot] += nd.coeff * v.imag();
                            }
                    }
                }
            }
        }
        const size_t seg_begin = batch.size();
        for (const RunOp &r : run) { // the run's U^dagger, in sweep order
            GateOp adj = *r.op;
            adj.inverse = !r.op->inverse;
            batch.push_back(std::move(adj));
        }
        run_segments.emplace_back(seg_begin, batch.size());
        run.clear();
    };

    for (long op_idx = static_cast<long>(ops.ops.size()) - 1; op_idx >= 0; op_idx--) {
        const GateOp &op = ops.ops[op_idx];
        if (op.name == "StatePrep" || op.name == "BasisState") // ADJ.hpp:447-450
            continue;
        if (tp_it == tp_rend) // ADJ.hpp:451-453
            break;
        if (runs_enabled) {
            RunOp r;
            r.op = &op;
            r.slot = -1;
            r.coeff = 0.0;
            if (one_qubit_matrix(op, &r.bit, r.u)) {
                bool ok = true;
                if (!op.params.empty()) {
                    if (current_param_idx == static_cast<long>(*tp_it)) {
                        double scale = 0.0;
                        ok = generator_2x2(op.name, r.g, &scale);
                        if (ok) {
                            r.slot = trainable_number;
                            r.coeff = -2.0 * scale * (op.inverse ? -1.0 : 1.0);
                            trainable_number--;
                            ++tp_it;
                        }
                    }
                    if (ok)
                        current_param_idx--;
                }
                if (ok) {
                    run.push_back(r);
                    continue;
                }
            }
            finish_run();
        }
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
    finish_run();
    // ops before the first trainable one never influence the Jacobian: the batch is dropped
    lambda->allreduce_device(d_jac, static_cast<int>(n_obs * tp_size));
    CUDA_CHECK(cudaMemcpyAsync(jac, d_jac, sizeof(double) * n_obs * tp_size, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < n_obs * tp_size; i++)
        jac[i] += jac_host[i];
    sv.last_adjoint_bytes = lambda->bytes_moved + (mu ? mu->bytes_moved : 0);
    for (const auto &h : H)
        sv.last_adjoint_bytes += h->bytes_moved;
    CUDA_CHECK(cudaFreeAsync(d_jac, st));
    if (d_tr_scratch) {
        CUDA_CHECK(cudaFreeAsync(d_tr_scratch, st));
        CUDA_CHECK(cudaFreeAsync(d_tr_out, st));
    }
}

} // namespace b2sv
