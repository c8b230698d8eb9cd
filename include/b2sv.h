/*
 * b2sv -- B200-native state-vector engine: the C ABI (the drop-in boundary).
 *
 * Every entry point below is what a binding of the reference's pybind11 module
 * `lightning_kokkos_qubit_ops` (reference pennylane_lightning_kokkos/src/bindings/Bindings.cpp)
 * would call instead of the Kokkos classes. Citations are reference file:line.
 *
 * Conventions (identical to the reference):
 *   - state = flat interleaved {re,im} array of 2^n amplitudes, complex64 or complex128
 *     (StateVectorKokkos.hpp:113); wire w <-> bit (n-1-w) of the flat index (GateFunctors.hpp:32);
 *   - for a k-wire gate wires[0] is the most-significant bit of the gate's local index,
 *     matrices are row-major (GateFunctors.hpp:73,96-97,175-189);
 *   - inverse != 0 applies U^dagger (GateFunctors.hpp:40-53,553-554).
 *   - gate parameters cross the ABI as double; complex host data that is not "the state"
 *     (matrices, CSR values, scattered amplitudes) crosses as interleaved complex128.
 *     State buffers (h2d/d2h) are in the state's own dtype.
 *
 * Ownership: host buffers are caller-owned and copied before the call returns; device
 * memory lives behind opaque handles. All functions return 0 on success, non-zero on error;
 * b2sv_last_error() gives the message (thread-local), formatted like the reference's
 * LightningException (Error.hpp:115-142). There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with an error.
 */
#ifndef B2SV_H
#define B2SV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2sv_state b2sv_state;   /* StateVectorKokkos<P>            SV.hpp:109      */
typedef struct b2sv_obs b2sv_obs;       /* ObservableKokkos<P> hierarchy   OBS.hpp:19-517  */
typedef struct b2sv_ops b2sv_ops;       /* OpsData<P>                      ADJ.hpp:17-173  */
typedef struct b2sv_csr b2sv_csr;       /* device-resident CSR matrix (new; ref re-uploads, MK.hpp:139-149) */

enum { B2SV_C64 = 0, B2SV_C128 = 1 };

/* ---- library ------------------------------------------------------------------------- */
const char *b2sv_last_error(void);
const char *b2sv_version(void);
/* replaces kokkos_config_info / print_configuration (Bindings.cpp:842-852) */
int b2sv_backend_info(char *buf, size_t cap);
int b2sv_device_count(int *count);

/* ---- state vector (SV.hpp:140-478 ctor, :492-532 init, :1596-1636 copies) ------------ */
int b2sv_create(int num_qubits, int dtype, int device_id, b2sv_state **out);
/* rank r of world 2^g holds the amplitudes whose top g index bits equal r (wires 0..g-1 global).
 * nccl_unique_id: 128 bytes from b2sv_comm_unique_id on rank 0, broadcast by the caller. */
int b2sv_create_sharded(int num_qubits_total, int dtype, int device_id, int rank, int world,
                        const void *nccl_unique_id, b2sv_state **out);
int b2sv_comm_unique_id(void *out128);
int b2sv_destroy(b2sv_state *s);
int b2sv_clone(const b2sv_state *src, b2sv_state **out);                 /* copy ctor SV.hpp:550-554 */
int b2sv_copy(b2sv_state *dst, const b2sv_state *src);                   /* updateData SV.hpp:1596 */
int b2sv_reset(b2sv_state *s);                                           /* resetStateVector :528 */
int b2sv_init_zeros(b2sv_state *s);                                      /* initZeros :483 */
int b2sv_set_basis_state(b2sv_state *s, uint64_t index);                 /* :492 */
int b2sv_set_state_vector(b2sv_state *s, const uint64_t *indices, const double *values_c128,
                          size_t n);                                     /* :503-521 */
/* state preparation on a subset of wires with the index table built on the device: all zeros, then
 * amplitude v of values_c128 (2^nw entries, wires[0] = MSB of v) lands on the basis state that has v
 * on `wires` and 0 elsewhere. Replaces the host-side itertools.product table of
 * lightning_kokkos.py:293-327 (_apply_state_vector_kokkos) + setStateVector. */
int b2sv_set_state_on_wires(b2sv_state *s, const int64_t *wires, int nw, const double *values_c128);
int b2sv_h2d(b2sv_state *s, const void *host, size_t length);            /* HostToDevice :1618 */
int b2sv_d2h(const b2sv_state *s, void *host, size_t length);            /* DeviceToHost :1626 */
/* sampled read (no reference counterpart; the reference copies the whole state, SV.hpp:1626):
 * out_c128[k] = amplitude at global flat index indices[k]; collective on sharded states */
int b2sv_get_amplitudes(const b2sv_state *s, const uint64_t *indices, size_t n, double *out_c128);
int b2sv_num_qubits(const b2sv_state *s, int *n);
int b2sv_data_length(const b2sv_state *s, uint64_t *len);                /* local length when sharded */
int b2sv_device_ptr(const b2sv_state *s, void **ptr);                    /* getData :1603 */
int b2sv_stream(const b2sv_state *s, void **cuda_stream);
int b2sv_sync(const b2sv_state *s);

/* ---- gates (applyOperation SV.hpp:585-600; applyOperation_std :611-628; lists :640-676) */
int b2sv_apply(b2sv_state *s, const char *name, const int64_t *wires, int nw, int inverse,
               const double *params, int np);
int b2sv_apply_matrix(b2sv_state *s, const int64_t *wires, int nw, int inverse,
                      const double *matrix_c128);
/* whole op list in one call: the fusion scheduler sees it all (Bindings.cpp:233-242) */
int b2sv_apply_ops(b2sv_state *s, const b2sv_ops *ops, int adjoint);
/* applyGenerator SV.hpp:687-695; returns the scaling factor through *scale */
int b2sv_apply_generator(b2sv_state *s, const char *name, const int64_t *wires, int nw, int adj,
                         double *scale);
/* fuse=0: one HBM sweep per gate (reference schedule); fuse=1 (default): tiled multi-gate passes */
int b2sv_set_fusion(b2sv_state *s, int fuse);
/* counters since the last reset: full-state sweeps executed, kernels launched */
int b2sv_get_stats(const b2sv_state *s, uint64_t *sweeps, uint64_t *launches);
int b2sv_reset_stats(b2sv_state *s);
/* algorithmic bytes the last b2sv_adjoint_jacobian / _vjp on this state moved over all its work
 * vectors (tile passes 2S, read passes S per vector read, copies 2S, Hamiltonian application) */
int b2sv_last_adjoint_traffic(const b2sv_state *s, uint64_t *bytes);
/* measurement aid: CUDA events around every tile pass (kind 0), generic-matrix kernel (1) and
 * global<->local exchange (2) launched on the state's stream between the two calls; trace_end
 * synchronises and returns up to `cap` records (*n = how many there were). Times in ms. */
int b2sv_trace_begin(b2sv_state *s);
int b2sv_trace_end(b2sv_state *s, int *kinds, double *start_ms, double *dur_ms, int cap, int *n);
/* developer aid: phase timers of the tile executor (all zero unless B2SV_TILE_PROF=1 is set in the
 * environment); reads and clears 16 cycle counters, see csrc/tile_kernel.cu g_tile_prof */
int b2sv_debug_tile_prof(uint64_t *out16);
/* sharded states: global<->local qubit swaps done so far, bytes each rank sent, and whether the
 * NVLink peer-memory swap kernel (1) or NCCL send/recv (0) carries them */
int b2sv_comm_stats(const b2sv_state *s, uint64_t *swaps, uint64_t *swap_bytes, int *peer_path);
/* bytes of pass descriptors / matrices the last apply call handed to the device */
int b2sv_last_upload_bytes(const b2sv_state *s, uint64_t *bytes);
/* sharded states keep swapped-in qubits where they are (lazy layout); this restores the identity
 * layout (wire w <-> bit n-1-w, top bits = rank). d2h does it implicitly. */
int b2sv_normalize_layout(b2sv_state *s);
/* current logical -> physical index-bit map (identity for single-GPU states); *n = number of bits */
int b2sv_layout(const b2sv_state *s, int *l2p, int cap, int *n);

/* ---- op lists (OpsData ADJ.hpp:40-56; create_ops_list Bindings.cpp:772-805) ---------- */
/* params / wires are concatenated; nparams[i] / nwires[i] give the split. matrices may be NULL;
 * otherwise matrices[i] is NULL or a row-major 2^k x 2^k complex128 matrix for op i. */
int b2sv_ops_create(int nops, const char *const *names, const double *params, const int *nparams,
                    const int64_t *wires, const int *nwires, const int *inverses,
                    const double *const *matrices_c128, b2sv_ops **out);
int b2sv_ops_destroy(b2sv_ops *ops);
int b2sv_ops_size(const b2sv_ops *ops, int *nops, int *n_par_ops);
/* Host-only (no device needed): what the fusion scheduler does with `ops` on an n-qubit state --
 * HBM passes, register rounds, arithmetic ops executed, permutation gates folded into the address
 * map for free, passes whose last round stores straight to HBM. The reference has no counterpart:
 * it runs one kernel per gate (StateVectorKokkos.hpp:807-824). */
int b2sv_plan_ops(const b2sv_ops *ops, int num_qubits, int dtype, uint64_t *passes,
                  uint64_t *rounds, uint64_t *arithmetic_ops, uint64_t *absorbed_perms,
                  uint64_t *fused_stores);

/* Host-only: the same for a state sharded over `world` ranks (no reference counterpart): runs of
 * shard-local work and global<->local exchanges. stats5: runs, exchanges, tile passes, exchanged bits,
 * bytes each rank sends (at 16 B per amplitude). buf (may be NULL) receives the plan as text. */
int b2sv_plan_sharded(const b2sv_ops *ops, int num_qubits, int world, int dtype, uint64_t *stats5,
                      char *buf, size_t cap);

/* ---- measurements (MeasuresKokkos.hpp) ----------------------------------------------- */
int b2sv_expval_named(const b2sv_state *s, const char *name, const int64_t *wires, int nw,
                      double *out);                                     /* MK.hpp:80-101,167-271 */
/* <Z_w> for every wire w in ONE read pass (the Python device asks for them one by one,
 * lightning_kokkos.py:554-559; b2sv_expval_named("PauliZ") is served from the same cached pass). */
int b2sv_expval_z_all(const b2sv_state *s, double *out, int cap);
/* drop cached measurements after writing through a pointer obtained from b2sv_device_ptr */
int b2sv_invalidate(b2sv_state *s);
int b2sv_expval_matrix(const b2sv_state *s, const int64_t *wires, int nw,
                       const double *matrix_c128, double *out);         /* MK.hpp:112-121,283-346 */
int b2sv_expval_csr(const b2sv_state *s, const double *data_c128, const uint64_t *indices,
                    const uint64_t *indptr, size_t nnz, size_t nrows, double *out); /* :132-157 */
int b2sv_csr_create(const b2sv_state *like, const double *data_c128, const uint64_t *indices,
                    const uint64_t *indptr, size_t nnz, size_t nrows, b2sv_csr **out);
int b2sv_csr_destroy(b2sv_csr *m);
int b2sv_expval_csr_resident(const b2sv_state *s, const b2sv_csr *m, double *out);
int b2sv_expval_obs(const b2sv_state *s, const b2sv_obs *ob, double *out);   /* MK.hpp:354-360 */
int b2sv_var_obs(const b2sv_state *s, const b2sv_obs *ob, double *out);      /* MK.hpp:368-381 */
/* wires==NULL / nw==0: all wires in order (MK.hpp:389-408); else marginal in the requested
 * wire order (MK.hpp:418-517). out has 2^nw doubles. */
int b2sv_probs(const b2sv_state *s, const int64_t *wires, int nw, double *out);
/* out: shots x num_qubits uint64, MSB (wire 0) first (MK.hpp:530-569, MF.hpp:113-115) */
int b2sv_generate_samples(const b2sv_state *s, size_t shots, uint64_t seed, uint64_t *out);
/* Re<a|b>, Im<a|b> (LinearAlgebraKokkos.hpp:155-236) and y += alpha x (:30-61) */
int b2sv_inner_product(const b2sv_state *a, const b2sv_state *b, double *re, double *im);
int b2sv_axpy(double alpha_re, double alpha_im, const b2sv_state *x, b2sv_state *y);

/* ---- observables (ObservablesKokkos.hpp; Bindings.cpp:591-736) ----------------------- */
int b2sv_obs_named(const char *name, const int64_t *wires, int nw, b2sv_obs **out);  /* OBS:77 */
int b2sv_obs_hermitian(const double *matrix_c128, const int64_t *wires, int nw,
                       b2sv_obs **out);                                              /* OBS:124 */
int b2sv_obs_tensor(b2sv_obs *const *obs, int n, b2sv_obs **out);                    /* OBS:185 */
int b2sv_obs_hamiltonian(const double *coeffs, b2sv_obs *const *obs, int n,
                         b2sv_obs **out);                                            /* OBS:296 */
int b2sv_obs_sparse(const double *data_c128, const uint64_t *indices, const uint64_t *indptr,
                    size_t nnz, size_t nrows, const int64_t *wires, int nw,
                    b2sv_obs **out);                                                 /* OBS:409 */
int b2sv_obs_destroy(b2sv_obs *ob);            /* children are reference-counted */
int b2sv_obs_name(const b2sv_obs *ob, char *buf, size_t cap);                        /* getObsName */
int b2sv_obs_wires(const b2sv_obs *ob, int64_t *wires, int cap, int *nw);            /* getWires */
int b2sv_obs_apply(const b2sv_obs *ob, b2sv_state *s);                               /* applyInPlace */

/* ---- adjoint Jacobian (AdjointDiffKokkos.hpp:404-478; Bindings.cpp:808-821) ---------- */
/* jac_out: row-major n_obs x n_tp doubles. trainable_params index the PARAMETRIC ops in order. */
int b2sv_adjoint_jacobian(const b2sv_state *s, b2sv_obs *const *obs, int n_obs,
                          const b2sv_ops *ops, const uint64_t *trainable_params, int n_tp,
                          double *jac_out);
/* vector-Jacobian product sum_o dy[o] * jac[o][p] (lightning_kokkos.py:689-727 vjp): computed as the
 * adjoint Jacobian of the single Hamiltonian sum_o dy[o] O_o, i.e. one reverse sweep. vjp_out: n_tp */
int b2sv_adjoint_vjp(const b2sv_state *s, b2sv_obs *const *obs, int n_obs, const double *dy,
                     const b2sv_ops *ops, const uint64_t *trainable_params, int n_tp,
                     double *vjp_out);

#ifdef __cplusplus
}
#endif
#endif /* B2SV_H */
