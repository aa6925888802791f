/*
 * gt_oracle.h -- CPU restatement ("oracle") of the GridTools stencil / gcl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the reported CPU baseline.  The product path
 * (gridtools_b200/, libgtb200.so) never links, imports or calls it.
 *
 * Parity pinning: every function here is checked in tests/test_oracle.py against
 *   (1) oracle/_ref/libgtref.so -- the reference's own cpu_ifirst / cpu_kfirst / naive backends
 *       compiled from /root/reference/include (see oracle/ref_driver.cpp, oracle/Makefile), and
 *   (2) the committed golden fixtures under tests/golden/ that were produced by that library
 *       together with the reference's analytic repositories
 *       (tests/regression/horizontal_diffusion_repository.hpp:32-79,
 *        tests/regression/vertical_advection_repository.hpp:70-151).
 *
 * Field convention (same as the product C-ABI, include/gtb200.h): a field is described by a pointer
 * to element (0,0,0) of the *compute domain* (first interior point, like the origin-shifted SIDs the
 * reference hands to a backend, stencil/core/backend.hpp:29-34) plus element strides (si, sj, sk).
 * Halo points are reached with negative / >= n indices.
 */
#ifndef GT_ORACLE_H
#define GT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    void *ptr;                         /* element (0,0,0) of the compute domain */
    int64_t stride_i, stride_j, stride_k; /* in elements */
} gto_field;

/* copy_stencil.cpp:24-36 -- out = in, bit exact.  elem_size in {4, 8}. */
int gto_copy(const gto_field *in, const gto_field *out, int ni, int nj, int nk, int elem_size);

/* horizontal_diffusion.cpp:35-106 -- lap / flx / fly / out, 4 stages, lap on [-1,1]^2 etc. */
int gto_hori_diff_f64(const gto_field *in, const gto_field *coeff, const gto_field *out, int ni, int nj, int nk);
int gto_hori_diff_f32(const gto_field *in, const gto_field *coeff, const gto_field *out, int ni, int nj, int nk);
/* simple_hori_diff.cpp:25-61; crlato / crlatu point at the compute domain's j = 0 and are read on [-1, nj] */
int gto_simple_hori_diff_f64(const gto_field *in, const gto_field *coeff, const double *crlato, const double *crlatu,
    const gto_field *out, int ni, int nj, int nk);
int gto_simple_hori_diff_f32(const gto_field *in, const gto_field *coeff, const float *crlato, const float *crlatu,
    const gto_field *out, int ni, int nj, int nk);

/* vertical_advection_dycore.cpp:32-149 -- forward elimination + back substitution, in place on utens_stage. */
int gto_vert_adv_f64(const gto_field *utens_stage, const gto_field *u_stage, const gto_field *wcon,
    const gto_field *u_pos, const gto_field *utens, double dtr_stage, int ni, int nj, int nk);
int gto_vert_adv_f32(const gto_field *utens_stage, const gto_field *u_stage, const gto_field *wcon,
    const gto_field *u_pos, const gto_field *utens, float dtr_stage, int ni, int nj, int nk);

/* tridiagonal.cpp:39-97 -- Thomas solve; sup and rhs are overwritten like in the reference. */
int gto_tridiagonal_f64(const gto_field *inf, const gto_field *diag, const gto_field *sup, const gto_field *rhs,
    const gto_field *out, int ni, int nj, int nk);

/* advection_pdbott_prepare_tracers.cpp:23-34 -- out[t] = rho * in[t] for n_tracers fields. */
int gto_prepare_tracers_f64(const gto_field *out, const gto_field *in, int n_tracers, const gto_field *rho,
    int ni, int nj, int nk);

/* ------------------------------------------------------------------------------------------------
 * gcl halo exchange (halo_exchange_dynamic_ut semantics), restated for R in-process "ranks".
 * common/halo_descriptor.hpp:44-227, gcl/high_level/descriptors.hpp:61-91 (pack / unpack loops),
 * gcl/low_level/proc_grids_3D.hpp:179-211 (neighbour lookup + periodicity),
 * gcl/high_level/descriptors_manual_gpu.hpp:176-253 (no clipping needed on the CPU path: messages to
 * non-existent neighbours are simply skipped, descriptors.hpp:577-579).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int minus, plus, begin, end, total; /* halo_descriptor(m, p, b, e, l), end inclusive */
} gto_halo;

/* Number of elements of the message sent to neighbour eta=(ei,ej,ek), each in {-1,0,1}:
 * prod_d s_length(eta_d)  (empty_field_base.hpp:170-176). */
int64_t gto_halo_send_count(const gto_halo h[3], int ei, int ej, int ek);
int64_t gto_halo_recv_count(const gto_halo h[3], int ei, int ej, int ek);

/* Pack the send region of one field for neighbour eta into buf (elem_size bytes per element), storage is
 * addressed as base[i + j*total0 + k*total0*total1] with i running over dimension 0 (descriptors.hpp:61-75).
 * Returns number of elements written. */
int64_t gto_halo_pack(const gto_halo h[3], int ei, int ej, int ek, const void *field, void *buf, int elem_size);
int64_t gto_halo_unpack(const gto_halo h[3], int ei, int ej, int ek, void *field, const void *buf, int elem_size);

/* Rank of the neighbour at offset (di,dj,dk) of process coordinates (pi,pj,pk) in a dims[3] Cartesian grid with
 * per-dimension periodicity; -1 if outside (proc_grids_3D.hpp:179-211).  Rank numbering is row-major like
 * MPI_Cart_create: rank = (pi*dims[1] + pj)*dims[2] + pk. */
int gto_proc_neighbour(const int dims[3], const int periodic[3], int pi, int pj, int pk, int di, int dj, int dk);

/* Full exchange among all ranks of a dims[3] process grid: fields[r*n_fields + f] is field f of rank r (all the
 * same halo layout h).  pack -> deliver -> unpack, exactly what pack()/exchange()/unpack() do together. */
/* boundaries/apply.hpp:44-56 with value_boundary / zero_boundary (kind 0) or copy_boundary (kind 1) */
int gto_boundary_apply(const gto_halo h[3], const int *mask, int kind, double value, void **fields, int n_fields,
    int elem_size);
int gto_halo_exchange_all(const gto_halo h[3], const int dims[3], const int periodic[3], void **fields, int n_fields,
    int elem_size);

#ifdef __cplusplus
}
#endif
#endif
