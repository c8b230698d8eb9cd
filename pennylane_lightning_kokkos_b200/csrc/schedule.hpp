// b2sv: fusion scheduler -- partitions a primitive-op list into HBM passes ("sweeps") and, inside a
// pass, into register rounds for the tile executor (tile_kernel.cu).
//
// A pass owns a tile of B index bits = the `low` contiguous low bits (coalescing) + (B - low) freely
// chosen bits. Every C1Q in the pass has its target among the tile bits; controls and DIAG masks may
// touch any bit (bits outside the tile are CTA-uniform predicates). Inside a pass each thread keeps
// 2^R amplitudes in registers; a round fixes which R tile bits are register-resident.
//
// Permutation gates are FREE inside a pass: the tile in shared memory carries a GF(2)-affine
// address map  slot(j) = phys(M j ^ o)  from the logical tile-local index j to the storage slot.
// A CNOT whose control and target are both tile bits multiplies M by an elementary matrix, an X
// (uncontrolled, or controlled only by bits outside the tile = CTA-uniform) toggles o; neither
// moves data. The scheduler folds them into the per-round address columns below; SWAP, CSWAP's
// and the CNOT conjugations of the Ising / excitation gates (gates.cpp) all vanish this way.
#pragma once
#include "ir.hpp"

namespace b2sv {

constexpr int kMaxRounds = 24;
constexpr int kMaxOpsPerPass = 80;
constexpr int kMaxTileBits = 16;
constexpr int kMaxRegBits = 5;
constexpr int kMaxFreeBits = 12; // tile bits that are not register bits (= log2 threads)
constexpr int kMaxCx = 32;       // conditional address toggles per pass
constexpr int kMaxDense = 6;
constexpr int kDefaultTileLow = 5; // 512-byte HBM runs (complex128)     // rounds 0 .. kMaxDense-1 of a pass may run in factored form

enum OpKind : uint8_t { KIND_GENERAL = 0, KIND_REAL = 1, KIND_PERM = 2, KIND_DIAG = 3 };
// flag bits stored in DevOp::kind above the OpKind
enum : uint8_t { OPF_UNCOND = 0x10, OPF_KIND_MASK = 0x0f }; // UNCOND: no control of any sort

// Device-side op record (read from shared memory by every thread; 16-byte aligned).
struct alignas(16) DevOp {
    double m[8];        // C1Q: m00,m01,m10,m11 as (re,im); DIAG: p0,p1 as (re,im)
    uint64_t gcm, gcv;  // control over bits outside the tile (CTA-uniform, tested on tile_base)
    uint64_t gpm;       // DIAG parity over bits outside the tile
    uint32_t lcm, lcv;  // control over tile-local, non-register LOGICAL bits (per thread)
    uint32_t lpm;       // DIAG parity over tile-local, non-register logical bits
    uint32_t slot_act;  // bit s: register slot s satisfies the register part of the control
    uint32_t slot_par;  // DIAG: parity contribution of register slot s
    uint8_t kind;       // OpKind | OPF_* flags
    uint8_t tslot;      // C1Q: which register bit (0..R-1) is the target
    int16_t jac;        // adjoint: Jacobian accumulator slot fed by this op, -1 = none
};
static_assert(sizeof(DevOp) == 112, "DevOp layout");

// Factored dense round (see schedule.cpp factor_2x2): every gate k of the round is written as
//   M_k = diag(P0_k, P1_k) * [[1, t01_k], [t10_k, 1]] * diag(1, q1_k)      (t real, |t| <~ 1)
// so the round is  post[s] * (prod_k shear_k) * pre[s]  with per-slot constants
//   pre[j] = prod_k q1_k^bit_k(j),  post[j] = prod_k P{bit_k(j)}_k,   j over the G gate slots:
// 2 FMAs per amplitude and gate + two complex multiplies per amplitude and round, instead of 8
// multiply-adds per amplitude and gate. Tables are stored in the state's own precision
// (16 x double2 or 32 x float2 = 256 bytes each).
struct alignas(16) DevDense {
    unsigned char pre[256];
    unsigned char post[256];
    unsigned char t[96]; // (t01_k, t10_k) per gate, in the state's real type
};
static_assert(sizeof(DevDense) == 608, "DevDense layout");

struct DevCx { // address toggle: from round `round` on, slot ^= vec when (tile_base & gcm) == gcv
    uint64_t gcm, gcv;
    uint16_t vec;   // phys(M e_t) at the time the X was absorbed
    uint16_t round; // first round (or n_rounds = the store phase) that sees it
    uint32_t pad_;
};

struct alignas(16) DevPassHeader {
    int32_t n_ops;
    int32_t n_rounds;
    int32_t low_bits;                 // number of contiguous low bits in the tile
    int32_t n_cx;
    uint8_t tile_bits[kMaxTileBits];  // ascending bit positions; tile_bits[j]=j for j<low_bits
    uint8_t round_regbits[kMaxRounds][8]; // logical tile-local positions held in registers, by slot
    // 0: generic round (op interpreter); 1 <= g <= 5: "dense" round = exactly g uncontrolled 2x2
    // gates, the k-th one on register slot k (straight-line code, no per-op dispatch);
    // 8 + g: the same in factored form, constants in PassParams::dense[rd] (indexed by the round, so
    // that the first rounds read them through uniform constant loads at compile-time offsets)
    uint8_t round_kind[kMaxRounds];
    uint8_t pad3_[kMaxRounds + 8];
    // tile id -> index with the tile bits cleared: base = OR_k ((id & seg_mask[k]) << seg_shift[k]),
    // one segment per run of consecutive non-tile index bits
    uint32_t seg_mask[kMaxTileBits + 1];
    uint8_t seg_shift[kMaxTileBits + 1];
    uint8_t n_seg;
    // how the tile leaves the SM: 0 = store phase (worker threads copy shared -> HBM through the
    // final address map); 1 = fused store, the last round writes its registers straight to HBM
    uint8_t fused_store;
    // 1: the tile lives in shared memory in plain order (slot = tile-local index) and is filled by bulk
    // async copies; 0: XOR-swizzled slots, filled amplitude by amplitude (cp.async)
    uint8_t plain_layout;
    uint8_t bulk_run_bits; // plain layout: the tile bits 0 .. bulk_run_bits-1 are the index bits of the same number
    uint8_t pad2_[11];
    // global index offsets (already pushed through the permutations absorbed after the last round):
    // of thread-id bit k of the last round, of register slot s, and of conditional toggle c
    uint64_t store_free[kMaxFreeBits];
    uint64_t store_reg[8];
    uint64_t store_cx[kMaxCx];
    // load warps: global BYTE offset of tile-local index e * 128 (the part of a load thread's
    // element address that does not depend on the thread), e < 2^B / 128
    uint64_t load_off[64];
    uint16_t round_begin[kMaxRounds + 1]; // op index ranges per round
    uint16_t pad_[3];
    // storage offset phys(M e_r) of register bit s in round rd
    uint16_t round_poff[kMaxRounds][8];
    // thread-id bit k <-> k-th non-register logical bit f_k: (1 << f_k) << 16 | phys(M e_{f_k})
    uint32_t round_col[kMaxRounds][kMaxFreeBits];
    // store phase: phys(M_final e_j) per logical tile-local bit j
    uint16_t final_col[kMaxTileBits];
    DevCx cx[kMaxCx];
};
static_assert(sizeof(DevPassHeader) % 16 == 0, "header must be copyable in 16-byte words");

// What one tile-kernel launch receives as its __grid_constant__ parameter.
struct PassParams {
    DevPassHeader hdr;
    DevOp ops[kMaxOpsPerPass];
    DevDense dense[kMaxDense];
};
static_assert(sizeof(PassParams) <= 32000, "kernel parameter space is 32764 bytes");

struct Pass {
    bool is_matk = false;
    Prim matk;              // when is_matk
    DevPassHeader hdr{};    // otherwise
    std::vector<DevOp> ops;
    std::vector<DevDense> dense; // factored dense rounds: entry rd belongs to round rd (others zero)
    std::vector<int> tags;  // per op: Prim::tag
    int n_absorbed = 0;     // permutation primitives folded into the address map
};

struct SchedConfig {
    int B = 12;        // tile bits
    int R = 4;         // register bits
    int SW = 3;        // log2(amplitudes per 128 B): the shared-memory swizzle width
    int SH = 0;        // log2(amplitudes per 16 B): index bits below SH are not swizzled (1 for complex64)
    int low = 5;       // contiguous low bits forced into every tile
    int n_local = 0;   // bits >= n_local cannot be targets (rank bits when sharded)
    int n_alloc = 0;   // index bits of the allocation (>= B; small states are zero-padded)
    bool fuse = true;  // false: one pass per primitive group (reference schedule)
    bool free_perms = true; // fold CNOT / X into the address map
    bool fuse_store = true; // let the last round of a pass write straight to HBM when coalescing allows
    // arithmetic ops per pass; 0 = automatic: build the schedule for several budgets and keep the
    // cheapest under the cost model of schedule_cost()
    int max_heavy = 0;
    bool f32 = false;       // complex64 state: factored-round tables are stored as float
    bool factor = true;     // factored dense rounds (D * shear * D form of the fused 2x2s)
    // grow the tile by marginal gain over a look-ahead window instead of first come, first served;
    // off by default: no fewer passes on the circuits tried, and it costs host time per pass
    bool lookahead = false;
    // allow plain-layout passes (bulk async tile loads) where the rounds permit it
    bool bulk = false;
    int bulk_min_run_bits = 8; // contiguous amplitudes per bulk copy: 2^8 x 16 B = 4 KiB (complex128)
};

// contiguous low index bits kept in every tile (2^low amplitudes per HBM run); B2SV_TILE_LOW overrides
int default_tile_low();

// Shared-memory swizzle (same function as tile_kernel.cuh phys<B,SW,SH>): XOR-folds every higher
// (SW - SH)-bit group of the index into bits SH .. SW-1. GF(2)-linear. SW = log2(amplitudes per 128 B),
// SH = log2(amplitudes per 16 B): complex64 keeps bit 0 in place, so the two amplitudes of a 16-byte
// unit stay together and the load warps copy 16 bytes at a time (L1-bypassing cp.async.cg).
inline uint32_t phys_slot(uint32_t i, int B, int SW, int SH = 0) {
    const int FW = SW - SH;
    uint32_t f = 0;
    for (int s = SH + FW; s < B; s += FW)
        f ^= (i >> s);
    return i ^ ((f & ((1u << FW) - 1u)) << SH);
}

// Pre-pass: merge runs of uncontrolled single-bit primitives on the same bit into one 2x2.
std::vector<Prim> fuse_single_qubit(const std::vector<Prim> &prims);
// Main entry.
std::vector<Pass> build_schedule(const std::vector<Prim> &prims, const SchedConfig &cfg);
// Tile id -> index of the tile's first amplitude: the id's bits are deposited into the index bits
// below `top` that are not in `excluded_mask` (the tile bits; for a pass over one slice of a shard
// also the bits that select the slice, whose values then come with the base pointer).
void fill_tile_id_segments(DevPassHeader &hdr, uint64_t excluded_mask, int top);
// Cost model (arbitrary units, fitted on 30-qubit complex128 sweeps on a B200): a pass costs what
// streaming the state costs, plus a smaller amount for every register round it runs on the tile.
double schedule_cost(const std::vector<Pass> &passes);

} // namespace b2sv
