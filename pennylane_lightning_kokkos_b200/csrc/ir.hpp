// b2sv: the primitive-op IR every gate, generator and Pauli observable is lowered to.
//
// Two fused primitives cover all 33 named gates and 19 generators of the reference
// (reference simulator/GateFunctors.hpp; formulas restated in SURVEY.md App. A):
//   C1Q  : a 2x2 complex matrix on ONE target bit, applied where (index & cmask) == cval
//   DIAG : amp *= (parity(index & pmask) ? p1 : p0), applied where (index & cmask) == cval
// Two-target pair rotations (IsingXX/XY/YY, SingleExcitation*, DoubleExcitation*, SWAP, CSWAP) are
// conjugated by CNOTs into a controlled C1Q (see gates.cpp), so the tile executor only ever needs
// single-target kernels. Arbitrary k>=2 qubit matrices stay as MATK (stand-alone kernel).
#pragma once
#include "common.hpp"

namespace b2sv {

struct Prim {
    enum Type : uint8_t { C1Q = 0, DIAG = 1, MATK = 2 };
    Type type = C1Q;
    int target = -1;        // C1Q: target bit position in the flat index
    uint64_t cmask = 0;     // control condition (index & cmask) == cval
    uint64_t cval = 0;
    uint64_t pmask = 0;     // DIAG: parity mask
    cplx m[4] = {1, 0, 0, 1}; // C1Q: row-major 2x2; DIAG: m[0]=p0 (even parity), m[1]=p1 (odd)
    // MATK only
    std::vector<int> bits;  // bit positions, bits[0] = MSB of the local index
    std::vector<cplx> mat;  // row-major 2^k x 2^k (already conjugate-transposed if inverse)
    // adjoint bookkeeping: >=0 marks the primitive that realises trainable op `jac_col`
    int tag = -1;

    uint64_t support() const { // every bit the primitive reads or writes
        uint64_t s = cmask | pmask;
        if (type == C1Q)
            s |= bit(target);
        for (int b : bits)
            s |= bit(b);
        return s;
    }
    uint64_t target_mask() const { // bits acted on non-diagonally
        if (type == C1Q)
            return bit(target);
        uint64_t s = 0;
        for (int b : bits)
            s |= bit(b);
        return s;
    }
};

// wires (reference convention, wire 0 = MSB) -> bit positions
inline std::vector<int> wires_to_bits(const std::vector<int64_t> &wires, int num_qubits) {
    std::vector<int> bits;
    bits.reserve(wires.size());
    for (auto w : wires) {
        B2_ABORT_IF(w < 0 || w >= num_qubits, "wire index out of range");
        bits.push_back(num_qubits - 1 - static_cast<int>(w));
    }
    for (size_t i = 0; i < bits.size(); i++)
        for (size_t j = i + 1; j < bits.size(); j++)
            B2_ABORT_IF(bits[i] == bits[j], "repeated wire in operation");
    return bits;
}

// Lower a named gate. Returns false when `name` is not a named gate of the reference
// (reference StateVectorKokkos.hpp:141-344 gates_ map); "Identity" lowers to nothing.
bool lower_gate(const std::string &name, const std::vector<int> &bits, bool inverse,
                const std::vector<double> &params, std::vector<Prim> &out);
// Lower a generator (reference StateVectorKokkos.hpp:345-461 generator_ map, wrappers :1275-1580).
// Returns false when no generator exists; *scale receives the reference's scaling factor.
bool lower_generator(const std::string &name, const std::vector<int> &bits, std::vector<Prim> &out,
                     double *scale);
// Generators that are plain Pauli words (RX, RY, RZ, IsingXX/YY/ZZ, MultiRZ): x / z masks over
// index bits and the number of Y factors, G|j> = i^ny (-1)^popc(j & z) |j ^ x>; scale as above.
// Lets the adjoint sweep take <H_lambda| G |lambda> in one read pass, without forming G|lambda>.
bool generator_pauli(const std::string &name, const std::vector<int> &bits, uint64_t *x,
                     uint64_t *z, int *ny, double *scale);
// Arbitrary matrix on `bits` (bits[0] = MSB of the local index), row-major; inverse = conj-transpose
// (reference GateFunctors.hpp:15-300, applyMultiQubitOp StateVectorKokkos.hpp:757-796).
void lower_matrix(const std::vector<int> &bits, bool inverse, const std::vector<cplx> &matrix,
                  std::vector<Prim> &out);
bool is_named_gate(const std::string &name);
int gate_num_params(const std::string &name); // -1 if unknown
int gate_num_wires(const std::string &name);  // 0 = any (MultiRZ), -1 if unknown

} // namespace b2sv
