// b2sv: lowering of the reference's named gates and generators to the C1Q / DIAG primitives.
//
// Semantics follow the reference functors (reference simulator/GateFunctors.hpp, line ranges cited
// per gate; formulas as restated in SURVEY.md App. A). Nothing here is a translation of those
// functors: each gate is re-derived as (a) a controlled single-target 2x2, (b) a parity/controlled
// phase, or (c) a CNOT-conjugated controlled 2x2 for the two-state "pair rotations"
//     PAIR(|..01..> <-> |..10..>)  =  CNOT(b->a) . C_a[M on b] . CNOT(b->a)
// which lets one single-target tile kernel execute every gate.
#include "ir.hpp"

#include <cmath>
#include <unordered_map>

namespace b2sv {
namespace {

const cplx I1{0.0, 1.0};

Prim c1q(int t, cplx m00, cplx m01, cplx m10, cplx m11, uint64_t cm = 0, uint64_t cv = 0) {
    Prim p;
    p.type = Prim::C1Q;
    p.target = t;
    p.cmask = cm;
    p.cval = cv;
    p.m[0] = m00;
    p.m[1] = m01;
    p.m[2] = m10;
    p.m[3] = m11;
    return p;
}
Prim diag(uint64_t pmask, cplx p0, cplx p1, uint64_t cm = 0, uint64_t cv = 0) {
    Prim p;
    p.type = Prim::DIAG;
    p.pmask = pmask;
    p.cmask = cm;
    p.cval = cv;
    p.m[0] = p0;
    p.m[1] = p1;
    p.m[2] = p.m[3] = 0;
    return p;
}
Prim xgate(int t, uint64_t cm = 0, uint64_t cv = 0) { return c1q(t, 0, 1, 1, 0, cm, cv); }
Prim ygate(int t, uint64_t cm = 0, uint64_t cv = 0) { return c1q(t, 0, -I1, I1, 0, cm, cv); }
Prim zgate(int t, uint64_t cm = 0, uint64_t cv = 0) { return diag(bit(t), 1, -1, cm, cv); }
Prim cnot(int c, int t, uint64_t extra = 0) { return xgate(t, bit(c) | extra, bit(c) | extra); }
Prim zero_where(uint64_t cm, uint64_t cv) { return diag(0, 0, 0, cm, cv); }

struct Mat2 {
    cplx a, b, c, d;
};
// 1-qubit rotation-family matrices; `inv` already folded in.
Mat2 m_rx(double th, bool inv) {
    const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
    return {c, -I1 * s, -I1 * s, c}; // GF.hpp:551-569
}
Mat2 m_ry(double th, bool inv) {
    const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
    return {c, -s, s, c}; // GF.hpp:590-608
}
Mat2 m_rot(double phi, double theta, double omega, bool inv) {
    if (inv) { // GF.hpp:3151-3163: (phi,theta,omega) -> (-omega,-theta,-phi)
        const double p = -omega, t = -theta, o = -phi;
        phi = p;
        theta = t;
        omega = o;
    }
    const double c = std::cos(theta / 2), s = std::sin(theta / 2);
    return {std::exp(-I1 * ((phi + omega) / 2)) * c, -std::exp(I1 * ((phi - omega) / 2)) * s,
            std::exp(-I1 * ((phi - omega) / 2)) * s, std::exp(I1 * ((phi + omega) / 2)) * c};
}
cplx phase(double ang, bool inv) { return std::exp(I1 * (inv ? -ang : ang)); }

struct Info {
    int nwires;  // 0 = variable
    int nparams;
};
const std::unordered_map<std::string, Info> &table() {
    static const std::unordered_map<std::string, Info> t = {
        {"PauliX", {1, 0}},
        {"PauliY", {1, 0}},
        {"PauliZ", {1, 0}},
        {"Hadamard", {1, 0}},
        {"S", {1, 0}},
        {"T", {1, 0}},
        {"PhaseShift", {1, 1}},
        {"RX", {1, 1}},
        {"RY", {1, 1}},
        {"RZ", {1, 1}},
        {"Rot", {1, 3}},
        {"CNOT", {2, 0}},
        {"CY", {2, 0}},
        {"CZ", {2, 0}},
        {"SWAP", {2, 0}},
        {"ControlledPhaseShift", {2, 1}},
        {"CRX", {2, 1}},
        {"CRY", {2, 1}},
        {"CRZ", {2, 1}},
        {"CRot", {2, 3}},
        {"IsingXX", {2, 1}},
        {"IsingXY", {2, 1}},
        {"IsingYY", {2, 1}},
        {"IsingZZ", {2, 1}},
        {"SingleExcitation", {2, 1}},
        {"SingleExcitationMinus", {2, 1}},
        {"SingleExcitationPlus", {2, 1}},
        {"DoubleExcitation", {4, 1}},
        {"DoubleExcitationMinus", {4, 1}},
        {"DoubleExcitationPlus", {4, 1}},
        {"MultiRZ", {0, 1}},
        {"CSWAP", {3, 0}},
        {"Toffoli", {3, 0}},
    };
    return t;
}

// PAIR rotation between |a=0,b=1> and |a=1,b=0>, expressed on x0 = v10, x1 = v01.
void pair_01_10(int a, int b, const Mat2 &m, std::vector<Prim> &out, uint64_t extra = 0) {
    out.push_back(cnot(b, a, extra));
    out.push_back(c1q(b, m.a, m.b, m.c, m.d, bit(a) | extra, bit(a) | extra));
    out.push_back(cnot(b, a, extra));
}
// PAIR rotation between |0011> and |1100> on bits (a,b,c,d), expressed on x0 = v1100, x1 = v0011.
void pair_0011_1100(const std::vector<int> &q, const Mat2 &m, std::vector<Prim> &out,
                    bool zero_rest = false) {
    const int a = q[0], b = q[1], c = q[2], d = q[3];
    out.push_back(cnot(d, a));
    out.push_back(cnot(d, b));
    out.push_back(cnot(d, c));
    if (zero_rest) { // generator of DoubleExcitation: everything outside the pair is annihilated
        out.push_back(zero_where(bit(a), 0));
        out.push_back(zero_where(bit(b), 0));
        out.push_back(zero_where(bit(c), bit(c)));
    }
    out.push_back(c1q(d, m.a, m.b, m.c, m.d, bit(a) | bit(b) | bit(c), bit(a) | bit(b)));
    out.push_back(cnot(d, c));
    out.push_back(cnot(d, b));
    out.push_back(cnot(d, a));
}
uint64_t mask_of(const std::vector<int> &bits) {
    uint64_t m = 0;
    for (int b : bits)
        m |= bit(b);
    return m;
}
} // namespace

bool is_named_gate(const std::string &name) { return table().count(name) != 0; }
int gate_num_params(const std::string &name) {
    auto it = table().find(name);
    return it == table().end() ? -1 : it->second.nparams;
}
int gate_num_wires(const std::string &name) {
    auto it = table().find(name);
    return it == table().end() ? -1 : it->second.nwires;
}

bool lower_gate(const std::string &name, const std::vector<int> &q, bool inv,
                const std::vector<double> &par, std::vector<Prim> &out) {
    if (name == "Identity")
        return true; // StateVectorKokkos.hpp:589-590: no-op
    auto it = table().find(name);
    if (it == table().end())
        return false;
    const Info info = it->second;
    B2_ABORT_IF(info.nwires != 0 && static_cast<int>(q.size()) != info.nwires,
                "Assertion failed: wires.size() == nqubits for gate " + name);
    B2_ABORT_IF(q.empty(), "gate " + name + " needs at least one wire");
    B2_ABORT_IF(static_cast<int>(par.size()) < info.nparams,
                "gate " + name + " needs " + std::to_string(info.nparams) + " parameter(s)");
    const double th = info.nparams >= 1 ? par[0] : 0.0;
    const double isq2 = 0.70710678118654752440;

    // ---- 1-qubit (GF.hpp:302-646, 3134-3186)
    if (name == "PauliX") {
        out.push_back(xgate(q[0]));
    } else if (name == "PauliY") {
        out.push_back(ygate(q[0]));
    } else if (name == "PauliZ") {
        out.push_back(zgate(q[0]));
    } else if (name == "Hadamard") {
        out.push_back(c1q(q[0], isq2, isq2, isq2, -isq2));
    } else if (name == "S") {
        out.push_back(diag(bit(q[0]), 1, inv ? -I1 : I1));
    } else if (name == "T") {
        out.push_back(diag(bit(q[0]), 1, phase(M_PI / 4, inv)));
    } else if (name == "PhaseShift") {
        out.push_back(diag(bit(q[0]), 1, phase(th, inv)));
    } else if (name == "RX") {
        const Mat2 m = m_rx(th, inv);
        out.push_back(c1q(q[0], m.a, m.b, m.c, m.d));
    } else if (name == "RY") {
        const Mat2 m = m_ry(th, inv);
        out.push_back(c1q(q[0], m.a, m.b, m.c, m.d));
    } else if (name == "RZ") {
        out.push_back(diag(bit(q[0]), phase(-th / 2, inv), phase(th / 2, inv)));
    } else if (name == "Rot") {
        const Mat2 m = m_rot(par[0], par[1], par[2], inv);
        out.push_back(c1q(q[0], m.a, m.b, m.c, m.d));
    }
    // ---- controlled 1-qubit, control = wires[0] (GF.hpp:649-848, 1769-1992)
    else if (name == "CNOT") {
        out.push_back(cnot(q[0], q[1]));
    } else if (name == "CY") {
        out.push_back(ygate(q[1], bit(q[0]), bit(q[0])));
    } else if (name == "CZ") {
        out.push_back(zgate(q[1], bit(q[0]), bit(q[0])));
    } else if (name == "ControlledPhaseShift") {
        out.push_back(diag(bit(q[1]), 1, phase(th, inv), bit(q[0]), bit(q[0])));
    } else if (name == "CRX") {
        const Mat2 m = m_rx(th, inv);
        out.push_back(c1q(q[1], m.a, m.b, m.c, m.d, bit(q[0]), bit(q[0])));
    } else if (name == "CRY") {
        const Mat2 m = m_ry(th, inv);
        out.push_back(c1q(q[1], m.a, m.b, m.c, m.d, bit(q[0]), bit(q[0])));
    } else if (name == "CRZ") {
        out.push_back(
            diag(bit(q[1]), phase(-th / 2, inv), phase(th / 2, inv), bit(q[0]), bit(q[0])));
    } else if (name == "CRot") {
        const Mat2 m = m_rot(par[0], par[1], par[2], inv);
        out.push_back(c1q(q[1], m.a, m.b, m.c, m.d, bit(q[0]), bit(q[0])));
    } else if (name == "Toffoli") { // GF.hpp:2063-2129: swap(v110, v111)
        out.push_back(xgate(q[2], bit(q[0]) | bit(q[1]), bit(q[0]) | bit(q[1])));
    }
    // ---- swaps (GF.hpp:851-892, 1995-2060)
    else if (name == "SWAP") {
        out.push_back(cnot(q[0], q[1]));
        out.push_back(cnot(q[1], q[0]));
        out.push_back(cnot(q[0], q[1]));
    } else if (name == "CSWAP") { // swap(v101, v110), control = wires[0]
        const uint64_t c = bit(q[0]);
        out.push_back(cnot(q[1], q[2], c));
        out.push_back(cnot(q[2], q[1], c));
        out.push_back(cnot(q[1], q[2], c));
    }
    // ---- Ising family (GF.hpp:895-1151); a = wires[0], b = wires[1]
    else if (name == "IsingXX") { // = CNOT(b->a) RX_b(theta) CNOT(b->a)
        const Mat2 m = m_rx(th, inv);
        out.push_back(cnot(q[1], q[0]));
        out.push_back(c1q(q[1], m.a, m.b, m.c, m.d));
        out.push_back(cnot(q[1], q[0]));
    } else if (name == "IsingYY") { // a=0 branch mixes (00,11) with +is, a=1 branch (10,01) with -is
        const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
        out.push_back(cnot(q[1], q[0]));
        out.push_back(c1q(q[1], c, I1 * s, I1 * s, c, bit(q[0]), 0));
        out.push_back(c1q(q[1], c, -I1 * s, -I1 * s, c, bit(q[0]), bit(q[0])));
        out.push_back(cnot(q[1], q[0]));
    } else if (name == "IsingXY") { // only (01,10) mix, with +is
        const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
        pair_01_10(q[0], q[1], {c, I1 * s, I1 * s, c}, out);
    } else if (name == "IsingZZ") {
        out.push_back(diag(bit(q[0]) | bit(q[1]), phase(-th / 2, inv), phase(th / 2, inv)));
    }
    // ---- excitations (GF.hpp:1154-1765)
    else if (name == "SingleExcitation" || name == "SingleExcitationMinus" ||
             name == "SingleExcitationPlus") {
        const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
        if (name == "SingleExcitationMinus")
            out.push_back(diag(bit(q[0]) | bit(q[1]), phase(-th / 2, inv), 1));
        if (name == "SingleExcitationPlus")
            out.push_back(diag(bit(q[0]) | bit(q[1]), phase(th / 2, inv), 1));
        // x0 = v10, x1 = v01:  x0' = c x0 + s x1,  x1' = -s x0 + c x1
        pair_01_10(q[0], q[1], {c, s, -s, c}, out);
    } else if (name == "DoubleExcitation" || name == "DoubleExcitationMinus" ||
               name == "DoubleExcitationPlus") {
        const double c = std::cos(th / 2), s = inv ? -std::sin(th / 2) : std::sin(th / 2);
        cplx ph = 1.0;
        if (name == "DoubleExcitationMinus")
            ph = phase(-th / 2, inv);
        if (name == "DoubleExcitationPlus")
            ph = phase(th / 2, inv);
        if (ph != cplx(1.0))
            out.push_back(diag(0, ph, ph)); // all 16 amps x ph, the pair gets M/ph below
        const cplx r = cplx(1.0) / ph;
        // x0 = v1100, x1 = v0011:  x0' = c x0 + s x1,  x1' = -s x0 + c x1
        pair_0011_1100(q, {c * r, s * r, -s * r, c * r}, out);
    } else if (name == "MultiRZ") { // GF.hpp:2132-2169
        out.push_back(diag(mask_of(q), phase(-th / 2, inv), phase(th / 2, inv)));
    } else {
        return false;
    }
    return true;
}

bool generator_pauli(const std::string &name, const std::vector<int> &q, uint64_t *x, uint64_t *z,
                     int *ny, double *scale) {
    *x = *z = 0;
    *ny = 0;
    *scale = -0.5;
    auto all = [&]() {
        uint64_t m = 0;
        for (int b : q)
            m |= bit(b);
        return m;
    };
    if (name == "RX" && q.size() == 1) {
        *x = bit(q[0]);
    } else if (name == "RY" && q.size() == 1) {
        *x = *z = bit(q[0]);
        *ny = 1;
    } else if (name == "RZ" && q.size() == 1) {
        *z = bit(q[0]);
    } else if (name == "IsingXX" && q.size() == 2) {
        *x = all();
    } else if (name == "IsingYY" && q.size() == 2) {
        *x = *z = all();
        *ny = 2;
    } else if (name == "IsingZZ" && q.size() == 2) {
        *z = all();
    } else if (name == "MultiRZ" && !q.empty()) {
        *z = all();
    } else {
        return false;
    }
    return true;
}

bool lower_generator(const std::string &name, const std::vector<int> &q, std::vector<Prim> &out,
                     double *scale) {
    auto need = [&](size_t n) {
        B2_ABORT_IF(q.size() != n, "Assertion failed: wires.size() == nqubits for generator " + name);
    };
    *scale = -0.5;
    if (name == "RX") { // SV.hpp:1452-1488: Pauli kernels
        need(1);
        out.push_back(xgate(q[0]));
    } else if (name == "RY") {
        need(1);
        out.push_back(ygate(q[0]));
    } else if (name == "RZ") {
        need(1);
        out.push_back(zgate(q[0]));
    } else if (name == "PhaseShift") { // GF.hpp:2192-2197: projector |1><1|
        need(1);
        out.push_back(diag(bit(q[0]), 0, 1));
        *scale = 1.0;
    } else if (name == "ControlledPhaseShift") { // GF.hpp:2948-2958: |11><11|
        need(2);
        out.push_back(zero_where(bit(q[0]), 0));
        out.push_back(zero_where(bit(q[1]), 0));
        *scale = 1.0;
    } else if (name == "CRX" || name == "CRY" || name == "CRZ") { // GF.hpp:2995-3107
        need(2);
        out.push_back(zero_where(bit(q[0]), 0));
        const uint64_t c = bit(q[0]);
        if (name == "CRX")
            out.push_back(xgate(q[1], c, c));
        else if (name == "CRY")
            out.push_back(ygate(q[1], c, c));
        else
            out.push_back(zgate(q[1], c, c));
    } else if (name == "IsingXX") { // GF.hpp:2235-2245  X(x)X
        need(2);
        out.push_back(xgate(q[0]));
        out.push_back(xgate(q[1]));
    } else if (name == "IsingYY") { // GF.hpp:2332-2344  Y(x)Y
        need(2);
        out.push_back(ygate(q[0]));
        out.push_back(ygate(q[1]));
    } else if (name == "IsingZZ") { // GF.hpp:2381-2390
        need(2);
        out.push_back(diag(bit(q[0]) | bit(q[1]), 1, -1));
    } else if (name == "IsingXY") { // GF.hpp:2283-2294: swap(10,01); 00,11 <- 0
        need(2);
        out.push_back(diag(bit(q[0]) | bit(q[1]), 0, 1));
        out.push_back(xgate(q[0]));
        out.push_back(xgate(q[1]));
        *scale = 0.5;
    } else if (name == "SingleExcitation") { // GF.hpp:2429-2443: Y_a X_b on the odd subspace
        need(2);
        out.push_back(diag(bit(q[0]) | bit(q[1]), 0, 1));
        out.push_back(ygate(q[0]));
        out.push_back(xgate(q[1]));
    } else if (name == "SingleExcitationMinus" || name == "SingleExcitationPlus") {
        need(2); // GF.hpp:2482-2547: v01' = -i v10, v10' = i v01 (x0 = v10, x1 = v01)
        if (name == "SingleExcitationPlus")
            out.push_back(diag(bit(q[0]) | bit(q[1]), -1, 1));
        pair_01_10(q[0], q[1], {0, I1, -I1, 0}, out);
    } else if (name == "DoubleExcitation") { // GF.hpp:2672-2689
        need(4);
        pair_0011_1100(q, {0, I1, -I1, 0}, out, /*zero_rest=*/true);
    } else if (name == "DoubleExcitationMinus") { // GF.hpp:2794-2799
        need(4);
        pair_0011_1100(q, {0, I1, -I1, 0}, out);
    } else if (name == "DoubleExcitationPlus") { // GF.hpp:2904-2909, scale SV.hpp:1436-1443
        need(4);
        pair_0011_1100(q, {0, -I1, I1, 0}, out);
        *scale = 0.5;
    } else if (name == "MultiRZ") { // GF.hpp:3127-3131
        B2_ABORT_IF(q.empty(), "MultiRZ generator needs wires");
        out.push_back(diag(mask_of(q), 1, -1));
    } else {
        return false;
    }
    return true;
}

void lower_matrix(const std::vector<int> &bits, bool inverse, const std::vector<cplx> &matrix,
                  std::vector<Prim> &out) {
    const size_t k = bits.size();
    B2_ABORT_IF(k == 0, "matrix operation needs at least one wire");
    B2_ABORT_IF(k > 10, "matrix operations on more than 10 wires are not supported");
    const size_t dim = size_t(1) << k;
    B2_ABORT_IF(matrix.size() != dim * dim, "matrix size does not match the number of wires");
    std::vector<cplx> m(dim * dim);
    for (size_t r = 0; r < dim; r++)
        for (size_t c = 0; c < dim; c++)
            m[r * dim + c] = inverse ? std::conj(matrix[c * dim + r]) : matrix[r * dim + c];
    if (k == 1) {
        out.push_back(c1q(bits[0], m[0], m[1], m[2], m[3]));
        return;
    }
    Prim p;
    p.type = Prim::MATK;
    p.bits = bits;
    p.mat = std::move(m);
    out.push_back(std::move(p));
}

} // namespace b2sv
