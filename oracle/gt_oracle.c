/*
 * gt_oracle.c -- plain-C CPU restatement of the GridTools stencil / gcl hot path.
 * TEST INFRASTRUCTURE ONLY (see gt_oracle.h).  Parity pinned: tests/test_oracle.py checks every function
 * against oracle/_ref/libgtref.so (the reference's own CPU backends compiled from /root/reference) and
 * against the golden fixtures in tests/golden/.
 *
 * Arithmetic is written with the same operand order as the reference's functors so that, compiled with
 * -ffp-contract=off, it is the "no-FMA" evaluation of exactly the reference expressions.  The reference
 * builds themselves (g++ -O3, nvcc) are free to contract a*b+c into FMAs, hence the 1e-12 / 1e-5
 * tolerances of the north star; pure copies are bit exact.
 */
#include "gt_oracle.h"

#include <stdlib.h>
#include <string.h>

#define AT(T, f, i, j, k) \
    (((T *)(f)->ptr)[(int64_t)(i) * (f)->stride_i + (int64_t)(j) * (f)->stride_j + (int64_t)(k) * (f)->stride_k])

/* ---------------------------------------------------------------- copy (copy_stencil.cpp:24-36) */
int gto_copy(const gto_field *in, const gto_field *out, int ni, int nj, int nk, int elem_size) {
    if (elem_size != 4 && elem_size != 8)
        return 1;
    for (int k = 0; k < nk; ++k)
        for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i) {
                if (elem_size == 8)
                    AT(uint64_t, out, i, j, k) = AT(uint64_t, in, i, j, k);
                else
                    AT(uint32_t, out, i, j, k) = AT(uint32_t, in, i, j, k);
            }
    return 0;
}

/* ------------------------------------------- horizontal diffusion (horizontal_diffusion.cpp:35-106)
 * Staged exactly like a CPU backend does it: lap on the (ni+2)x(nj+2) region, flx on (ni+1)xnj,
 * fly on nix(nj+1), then out (extents from SURVEY Appendix A / compute_extents). */
#define GTO_HORI_DIFF(NAME, T)                                                                                   \
    int NAME(const gto_field *in, const gto_field *coeff, const gto_field *out, int ni, int nj, int nk) {        \
        const int li = ni + 2, lj = nj + 2;                                                                      \
        T *lap = (T *)malloc(sizeof(T) * (size_t)li * lj);                                                       \
        T *flx = (T *)malloc(sizeof(T) * (size_t)li * lj);                                                       \
        T *fly = (T *)malloc(sizeof(T) * (size_t)li * lj);                                                       \
        if (!lap || !flx || !fly) {                                                                              \
            free(lap), free(flx), free(fly);                                                                     \
            return 2;                                                                                            \
        }                                                                                                        \
        for (int k = 0; k < nk; ++k) {                                                                           \
            /* lap_function :35-47 : 4*in - (in(1,0) + in(0,1) + in(-1,0) + in(0,-1)) */                         \
            for (int j = -1; j <= nj; ++j)                                                                       \
                for (int i = -1; i <= ni; ++i)                                                                   \
                    lap[(j + 1) * li + (i + 1)] =                                                                \
                        (T)4 * AT(T, in, i, j, k) - (AT(T, in, i + 1, j, k) + AT(T, in, i, j + 1, k) +            \
                                                        AT(T, in, i - 1, j, k) + AT(T, in, i, j - 1, k));        \
            /* flx_function :49-61 */                                                                            \
            for (int j = 0; j < nj; ++j)                                                                         \
                for (int i = -1; i < ni; ++i) {                                                                  \
                    T res = lap[(j + 1) * li + (i + 2)] - lap[(j + 1) * li + (i + 1)];                           \
                    flx[(j + 1) * li + (i + 1)] =                                                                \
                        res * (AT(T, in, i + 1, j, k) - AT(T, in, i, j, k)) > 0 ? (T)0 : res;                    \
                }                                                                                                \
            /* fly_function :63-75 */                                                                            \
            for (int j = -1; j < nj; ++j)                                                                        \
                for (int i = 0; i < ni; ++i) {                                                                   \
                    T res = lap[(j + 2) * li + (i + 1)] - lap[(j + 1) * li + (i + 1)];                           \
                    fly[(j + 1) * li + (i + 1)] =                                                                \
                        res * (AT(T, in, i, j + 1, k) - AT(T, in, i, j, k)) > 0 ? (T)0 : res;                    \
                }                                                                                                \
            /* out_function :77-91 : in - coeff * (flx - flx(-1,0) + fly - fly(0,-1)) */                         \
            for (int j = 0; j < nj; ++j)                                                                         \
                for (int i = 0; i < ni; ++i)                                                                     \
                    AT(T, out, i, j, k) =                                                                        \
                        AT(T, in, i, j, k) -                                                                     \
                        AT(T, coeff, i, j, k) * (flx[(j + 1) * li + (i + 1)] - flx[(j + 1) * li + i] +           \
                                                    fly[(j + 1) * li + (i + 1)] - fly[j * li + (i + 1)]);        \
        }                                                                                                        \
        free(lap), free(flx), free(fly);                                                                         \
        return 0;                                                                                                \
    }

GTO_HORI_DIFF(gto_hori_diff_f64, double)
GTO_HORI_DIFF(gto_hori_diff_f32, float)

/* --------------------------------- simple horizontal diffusion (simple_hori_diff.cpp:25-61)
 * crlato / crlatu depend on j only (selector<0,1,0> storages in the reference): plain arrays indexed by the compute
 * domain's j, valid on [-1, nj].  wlap_function :25-41 on the 1-extended domain, divflux_function :43-61. */
#define GTO_SIMPLE_HORI_DIFF(NAME, T)                                                                            \
    int NAME(const gto_field *in, const gto_field *coeff, const T *crlato, const T *crlatu, const gto_field *out, \
        int ni, int nj, int nk) {                                                                                \
        const int li = ni + 2, lj = nj + 2;                                                                      \
        T *lap = (T *)malloc(sizeof(T) * (size_t)li * lj);                                                       \
        if (!lap)                                                                                                \
            return 2;                                                                                            \
        for (int k = 0; k < nk; ++k) {                                                                           \
            for (int j = -1; j <= nj; ++j)                                                                       \
                for (int i = -1; i <= ni; ++i)                                                                   \
                    lap[(j + 1) * li + (i + 1)] =                                                                \
                        AT(T, in, i + 1, j, k) + AT(T, in, i - 1, j, k) - (T)2 * AT(T, in, i, j, k) +            \
                        crlato[j] * (AT(T, in, i, j + 1, k) - AT(T, in, i, j, k)) +                              \
                        crlatu[j] * (AT(T, in, i, j - 1, k) - AT(T, in, i, j, k));                               \
            for (int j = 0; j < nj; ++j)                                                                         \
                for (int i = 0; i < ni; ++i) {                                                                   \
                    const T c = lap[(j + 1) * li + (i + 1)];                                                     \
                    T fluxx = lap[(j + 1) * li + (i + 2)] - c;                                                   \
                    T fluxx_m = c - lap[(j + 1) * li + i];                                                       \
                    T fluxy = crlato[j] * (lap[(j + 2) * li + (i + 1)] - c);                                     \
                    T fluxy_m = crlato[j] * (c - lap[j * li + (i + 1)]);                                         \
                    AT(T, out, i, j, k) =                                                                        \
                        AT(T, in, i, j, k) + ((fluxx_m - fluxx) + (fluxy_m - fluxy)) * AT(T, coeff, i, j, k);    \
                }                                                                                                \
        }                                                                                                        \
        free(lap);                                                                                               \
        return 0;                                                                                                \
    }

GTO_SIMPLE_HORI_DIFF(gto_simple_hori_diff_f64, double)
GTO_SIMPLE_HORI_DIFF(gto_simple_hori_diff_f32, float)

/* --------------------------------- vertical advection (vertical_advection_dycore.cpp:32-149)
 * BET_M = BET_P = 0.5 (vertical_advection_defs.hpp).  Forward sweep: first_level :85-98, body :50-68,
 * last_level :70-83.  Backward sweep: last_level :118-121, body :111-116.  ccol/dcol are column
 * temporaries; data_col is the running back-substituted value. */
#define GTO_VERT_ADV(NAME, T)                                                                                      \
    int NAME(const gto_field *utens_stage, const gto_field *u_stage, const gto_field *wcon, const gto_field *u_pos, \
        const gto_field *utens, T dtr_stage, int ni, int nj, int nk) {                                             \
        const T bet_m = (T)0.5, bet_p = (T)0.5;                                                                    \
        if (nk < 2)                                                                                                \
            return 1;                                                                                              \
        T *ccol = (T *)malloc(sizeof(T) * (size_t)nk);                                                             \
        T *dcol = (T *)malloc(sizeof(T) * (size_t)nk);                                                             \
        if (!ccol || !dcol) {                                                                                      \
            free(ccol), free(dcol);                                                                                \
            return 2;                                                                                              \
        }                                                                                                          \
        for (int j = 0; j < nj; ++j)                                                                               \
            for (int i = 0; i < ni; ++i) {                                                                         \
                for (int k = 0; k < nk; ++k) {                                                                     \
                    T dd = dtr_stage * AT(T, u_pos, i, j, k) + AT(T, utens, i, j, k) + AT(T, utens_stage, i, j, k); \
                    if (k == 0) {                                                                                  \
                        T gcv = (T).25 * (AT(T, wcon, i + 1, j, k + 1) + AT(T, wcon, i, j, k + 1));                \
                        T cs = gcv * bet_m;                                                                        \
                        T c = gcv * bet_p;                                                                         \
                        T b = dtr_stage - c;                                                                       \
                        T correction = -cs * (AT(T, u_stage, i, j, k + 1) - AT(T, u_stage, i, j, k));              \
                        T d = dd + correction;                                                                     \
                        T divided = (T)1 / b;                                                                      \
                        ccol[k] = c * divided;                                                                     \
                        dcol[k] = d * divided;                                                                     \
                    } else if (k < nk - 1) {                                                                       \
                        T gav = -(T).25 * (AT(T, wcon, i + 1, j, k) + AT(T, wcon, i, j, k));                       \
                        T gcv = (T).25 * (AT(T, wcon, i + 1, j, k + 1) + AT(T, wcon, i, j, k + 1));                \
                        T as = gav * bet_m;                                                                        \
                        T cs = gcv * bet_m;                                                                        \
                        T a = gav * bet_p;                                                                         \
                        T c = gcv * bet_p;                                                                         \
                        T b = dtr_stage - a - c;                                                                   \
                        T correction = -as * (AT(T, u_stage, i, j, k - 1) - AT(T, u_stage, i, j, k)) -             \
                                       cs * (AT(T, u_stage, i, j, k + 1) - AT(T, u_stage, i, j, k));               \
                        T d = dd + correction;                                                                     \
                        T divided = (T)1 / (b - ccol[k - 1] * a);                                                  \
                        ccol[k] = c * divided;                                                                     \
                        dcol[k] = (d - dcol[k - 1] * a) * divided;                                                 \
                    } else {                                                                                       \
                        T gav = -(T).25 * (AT(T, wcon, i + 1, j, k) + AT(T, wcon, i, j, k));                       \
                        T as = gav * bet_m;                                                                        \
                        T a = gav * bet_p;                                                                         \
                        T b = dtr_stage - a;                                                                       \
                        T correction = -as * (AT(T, u_stage, i, j, k - 1) - AT(T, u_stage, i, j, k));              \
                        T d = dd + correction;                                                                     \
                        T divided = (T)1 / (b - ccol[k - 1] * a);                                                  \
                        dcol[k] = (d - dcol[k - 1] * a) * divided;                                                 \
                    }                                                                                              \
                }                                                                                                  \
                T data = dcol[nk - 1];                                                                             \
                AT(T, utens_stage, i, j, nk - 1) = dtr_stage * (data - AT(T, u_pos, i, j, nk - 1));                \
                for (int k = nk - 2; k >= 0; --k) {                                                                \
                    data = dcol[k] - ccol[k] * data;                                                               \
                    AT(T, utens_stage, i, j, k) = dtr_stage * (data - AT(T, u_pos, i, j, k));                      \
                }                                                                                                  \
            }                                                                                                      \
        free(ccol), free(dcol);                                                                                    \
        return 0;                                                                                                  \
    }

GTO_VERT_ADV(gto_vert_adv_f64, double)
GTO_VERT_ADV(gto_vert_adv_f32, float)

/* ----------------------------------------------- Thomas solve (tridiagonal.cpp:39-74)
 * forward :46-56 (k >= 1) and :58-62 (k == 0); backward :71-74 and :65-69. */
int gto_tridiagonal_f64(const gto_field *inf, const gto_field *diag, const gto_field *sup, const gto_field *rhs,
    const gto_field *out, int ni, int nj, int nk) {
    typedef double T;
    for (int j = 0; j < nj; ++j)
        for (int i = 0; i < ni; ++i) {
            AT(T, sup, i, j, 0) = AT(T, sup, i, j, 0) / AT(T, diag, i, j, 0);
            AT(T, rhs, i, j, 0) = AT(T, rhs, i, j, 0) / AT(T, diag, i, j, 0);
            for (int k = 1; k < nk; ++k) {
                T den = AT(T, diag, i, j, k) - AT(T, sup, i, j, k - 1) * AT(T, inf, i, j, k);
                AT(T, sup, i, j, k) = AT(T, sup, i, j, k) / den;
                AT(T, rhs, i, j, k) = (AT(T, rhs, i, j, k) - AT(T, inf, i, j, k) * AT(T, rhs, i, j, k - 1)) / den;
            }
            AT(T, out, i, j, nk - 1) = AT(T, rhs, i, j, nk - 1);
            for (int k = nk - 2; k >= 0; --k)
                AT(T, out, i, j, k) = AT(T, rhs, i, j, k) - AT(T, sup, i, j, k) * AT(T, out, i, j, k + 1);
        }
    return 0;
}

/* ------------------------------ prepare_tracers (advection_pdbott_prepare_tracers.cpp:23-34) */
int gto_prepare_tracers_f64(const gto_field *out, const gto_field *in, int n_tracers, const gto_field *rho, int ni,
    int nj, int nk) {
    for (int t = 0; t < n_tracers; ++t)
        for (int k = 0; k < nk; ++k)
            for (int j = 0; j < nj; ++j)
                for (int i = 0; i < ni; ++i)
                    AT(double, &out[t], i, j, k) = AT(double, rho, i, j, k) * AT(double, &in[t], i, j, k);
    return 0;
}

/* ===================================================================== gcl halo exchange ========= */

/* common/halo_descriptor.hpp:90-164 */
static int lo_inside(const gto_halo *h, int e) { return e == 1 ? h->end - h->minus + 1 : h->begin; }
static int hi_inside(const gto_halo *h, int e) { return e == -1 ? h->begin + h->plus - 1 : h->end; }
static int lo_outside(const gto_halo *h, int e) {
    return e == 0 ? h->begin : (e == 1 ? h->end + 1 : h->begin - h->minus);
}
static int hi_outside(const gto_halo *h, int e) { return e == 0 ? h->end : (e == 1 ? h->end + h->plus : h->begin - 1); }
/* :166-201 */
static int s_length(const gto_halo *h, int e) { return e == 0 ? h->end - h->begin + 1 : (e == -1 ? h->plus : h->minus); }
static int r_length(const gto_halo *h, int e) { return e == 0 ? h->end - h->begin + 1 : (e == 1 ? h->plus : h->minus); }

int64_t gto_halo_send_count(const gto_halo h[3], int ei, int ej, int ek) {
    return (int64_t)s_length(&h[0], ei) * s_length(&h[1], ej) * s_length(&h[2], ek);
}
int64_t gto_halo_recv_count(const gto_halo h[3], int ei, int ej, int ek) {
    return (int64_t)r_length(&h[0], ei) * r_length(&h[1], ej) * r_length(&h[2], ek);
}

/* gcl/high_level/descriptors.hpp:61-75 */
int64_t gto_halo_pack(const gto_halo h[3], int ei, int ej, int ek, const void *field, void *buf, int elem_size) {
    const char *src = (const char *)field;
    char *dst = (char *)buf;
    int64_t n = 0;
    for (int k = lo_inside(&h[2], ek); k <= hi_inside(&h[2], ek); ++k)
        for (int j = lo_inside(&h[1], ej); j <= hi_inside(&h[1], ej); ++j)
            for (int i = lo_inside(&h[0], ei); i <= hi_inside(&h[0], ei); ++i) {
                int64_t idx = i + (int64_t)h[0].total * (j + (int64_t)h[1].total * k);
                memcpy(dst + n * elem_size, src + idx * elem_size, (size_t)elem_size);
                ++n;
            }
    return n;
}

/* gcl/high_level/descriptors.hpp:77-91 */
int64_t gto_halo_unpack(const gto_halo h[3], int ei, int ej, int ek, void *field, const void *buf, int elem_size) {
    char *dst = (char *)field;
    const char *src = (const char *)buf;
    int64_t n = 0;
    for (int k = lo_outside(&h[2], ek); k <= hi_outside(&h[2], ek); ++k)
        for (int j = lo_outside(&h[1], ej); j <= hi_outside(&h[1], ej); ++j)
            for (int i = lo_outside(&h[0], ei); i <= hi_outside(&h[0], ei); ++i) {
                int64_t idx = i + (int64_t)h[0].total * (j + (int64_t)h[1].total * k);
                memcpy(dst + idx * elem_size, src + n * elem_size, (size_t)elem_size);
                ++n;
            }
    return n;
}

/* gcl/low_level/proc_grids_3D.hpp:179-211 */
int gto_proc_neighbour(const int dims[3], const int periodic[3], int pi, int pj, int pk, int di, int dj, int dk) {
    int c[3] = {pi + di, pj + dj, pk + dk};
    for (int d = 0; d < 3; ++d) {
        if (periodic[d])
            c[d] = (c[d] + dims[d]) % dims[d];
        else if (c[d] < 0 || c[d] >= dims[d])
            return -1;
    }
    return (c[0] * dims[1] + c[1]) * dims[2] + c[2];
}

/* pack -> exchange -> unpack for every rank of the process grid (descriptors.hpp:559-624 plus the
 * Isend/Irecv pairing of low_level/Halo_Exchange_3D.hpp:551-790: what rank r sends towards eta is what the
 * neighbour receives from -eta). */
int gto_halo_exchange_all(const gto_halo h[3], const int dims[3], const int periodic[3], void **fields, int n_fields,
    int elem_size) {
    const int nranks = dims[0] * dims[1] * dims[2];
    int64_t maxcount = 0;
    for (int a = -1; a <= 1; ++a)
        for (int b = -1; b <= 1; ++b)
            for (int c = -1; c <= 1; ++c) {
                int64_t n = gto_halo_send_count(h, a, b, c);
                if (n > maxcount)
                    maxcount = n;
            }
    /* all messages are packed before any is unpacked, like pack() ; exchange() ; unpack() */
    char *bufs = (char *)malloc((size_t)nranks * 27 * (size_t)n_fields * (size_t)maxcount * (size_t)elem_size + 1);
    if (!bufs)
        return 2;
    const size_t slot = (size_t)n_fields * (size_t)maxcount * (size_t)elem_size;
    for (int r = 0; r < nranks; ++r) {
        int pk = r % dims[2], pj = (r / dims[2]) % dims[1], pi = r / (dims[2] * dims[1]);
        for (int a = -1; a <= 1; ++a)
            for (int b = -1; b <= 1; ++b)
                for (int c = -1; c <= 1; ++c) {
                    if (!a && !b && !c)
                        continue;
                    if (gto_proc_neighbour(dims, periodic, pi, pj, pk, a, b, c) < 0)
                        continue;
                    char *buf = bufs + ((size_t)r * 27 + (size_t)((c + 1) * 9 + (b + 1) * 3 + (a + 1))) * slot;
                    for (int f = 0; f < n_fields; ++f)
                        buf += elem_size * gto_halo_pack(h, a, b, c, fields[r * n_fields + f], buf, elem_size);
                }
    }
    for (int r = 0; r < nranks; ++r) {
        int pk = r % dims[2], pj = (r / dims[2]) % dims[1], pi = r / (dims[2] * dims[1]);
        for (int a = -1; a <= 1; ++a)
            for (int b = -1; b <= 1; ++b)
                for (int c = -1; c <= 1; ++c) {
                    if (!a && !b && !c)
                        continue;
                    int q = gto_proc_neighbour(dims, periodic, pi, pj, pk, a, b, c);
                    if (q < 0)
                        continue;
                    /* r receives from q what q sent towards -eta */
                    const char *buf =
                        bufs + ((size_t)q * 27 + (size_t)((-c + 1) * 9 + (-b + 1) * 3 + (-a + 1))) * slot;
                    for (int f = 0; f < n_fields; ++f)
                        buf += elem_size * gto_halo_unpack(h, a, b, c, fields[r * n_fields + f], buf, elem_size);
                }
    }
    free(bufs);
    return 0;
}

/* --------------------------------- boundary conditions (boundaries/apply.hpp:44-56, value.hpp:27-66, zero.hpp,
 * copy.hpp:26-46).  For every direction (ei, ej, ek) != 0 with mask[n] != 0 (n = (ei+1) + 3 (ej+1) + 9 (ek+1); mask
 * NULL = default_predicate) the functor runs on loop_{low,high}_bound_outside of the three halo descriptors.
 * kind 0: every field = value; kind 1: fields[0 .. n-2] = fields[n-1] (copy_boundary). */
int gto_boundary_apply(const gto_halo h[3], const int *mask, int kind, double value, void **fields, int n_fields,
    int elem_size) {
    const int n_dst = kind == 1 ? n_fields - 1 : n_fields;
    if (n_dst < 1 || (elem_size != 4 && elem_size != 8) || (kind != 0 && kind != 1))
        return 1;
    const int64_t s1 = h[0].total, s2 = (int64_t)h[0].total * h[1].total;
    for (int ek = -1; ek <= 1; ++ek)
        for (int ej = -1; ej <= 1; ++ej)
            for (int ei = -1; ei <= 1; ++ei) {
                const int n = (ei + 1) + 3 * (ej + 1) + 9 * (ek + 1);
                if (n == 13 || (mask && !mask[n]))
                    continue;
                for (int k = lo_outside(&h[2], ek); k <= hi_outside(&h[2], ek); ++k)
                    for (int j = lo_outside(&h[1], ej); j <= hi_outside(&h[1], ej); ++j)
                        for (int i = lo_outside(&h[0], ei); i <= hi_outside(&h[0], ei); ++i) {
                            const int64_t o = i + j * s1 + k * s2;
                            for (int f = 0; f < n_dst; ++f) {
                                if (elem_size == 8)
                                    ((double *)fields[f])[o] = kind == 1 ? ((const double *)fields[n_fields - 1])[o] : value;
                                else
                                    ((float *)fields[f])[o] =
                                        kind == 1 ? ((const float *)fields[n_fields - 1])[o] : (float)value;
                            }
                        }
            }
    return 0;
}
