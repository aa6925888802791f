// vert_adv.cu -- vertical sweeps for sm_100a: vertical_advection_dycore (vertical_advection_dycore.cpp:32-149)
// and the plain Thomas solve (tridiagonal.cpp:39-97).
//
// What the reference does (stencil/gpu/make_kernel_fun.hpp:50-98, k_cache.hpp:24-46, fill_flush.hpp:123-324): one
// thread per column, forward and backward MSS fused into one launch, u_stage in a 3-deep register k_cache filled
// level by level, ccol/dcol in 2-deep register k_caches that are flushed to a blocked global temporary and read back
// by the backward sweep; no prefetching ("no unrolling", make_kernel_fun.hpp:62-64), so every level waits for its
// own loads.
//
// What this kernel does instead:
//  * one thread per column, a warp covers 32 consecutive i (256 B per fp64 load instruction), CTAs are small
//    (default 64 threads) so that 256x256 columns spread evenly over 148 SMs;
//  * the k_caches are plain registers rotated by the sweep: u_stage(k-1,k,k+1), wcon(i,k)+wcon(i+1,k) (the sum is
//    what both gav of level k and gcv of level k-1 need), ccol/dcol(k-1), data_col(k+1);
//  * loads are software-pipelined UNROLL levels ahead through a register ring (all loads of the next UNROLL levels
//    are issued before the current UNROLL levels are computed), which is what hides HBM latency when only ~450
//    columns live on an SM;
//  * ccol/dcol, which the reference flushes to HBM, go either to shared memory (k-major, conflict free) when
//    2*nk*threads elements fit, or to a column-interleaved global scratch accessed with an L2 evict_last policy
//    while the streamed fields use evict_first, so the flush/read-back stays on chip as far as L2 allows.
//
// Arithmetic follows the functor bodies operation by operation (-fmad=false, IEEE division), so results are
// bit-identical to oracle/gt_oracle.c compiled with -ffp-contract=off.
#include "common.cuh"

using namespace gtb;

namespace {

    template <class T>
    struct col_field {
        T *ptr;
        int64_t sj, sk;
    };

    template <class T>
    struct va_params {
        col_field<T> utens_stage;
        col_field<const T> u_stage, wcon, u_pos, utens;
        T dtr;
        int ni, nj, nk;
        int tiles_i;
        int items;      // column blocks of 32 x (threads/32) columns
        T *scratch;     // [NS*nk][slots] when the global scratch is used
        int64_t slots;  // columns (one slot per column) or resident threads (persistent grid: one slot per thread)
        int persistent; // slots are per thread and CTAs loop over the items
    };

    template <class T, bool Hints>
    __device__ __forceinline__ T ldg_stream(const T *p, uint64_t pol) {
        if constexpr (Hints)
            return ptx::ld_hint(p, pol);
        else
            return __ldg(p);
    }

    template <class T>
    struct va_level {
        T us, un, w0, w1, up, ut; // utens_stage(k), u_stage(k+1), wcon(i,k+1), wcon(i+1,k+1), u_pos(k), utens(k)
    };

    // SMEM: ccol/dcol live in dynamic shared memory [NS*nk][THREADS]; else in p.scratch [NS*nk][slots].
    // SAVE_UPOS: u_pos(k) is kept next to ccol/dcol (NS = 3) so that the backward sweep does not read it from HBM a
    // second time.
    template <class T, int UNROLL, bool SMEM, bool Hints, bool SAVE_UPOS>
    __device__ __forceinline__ void va_columns(const va_params<T> &p, int item, unsigned char *smem_raw) {
        constexpr int NS = SAVE_UPOS ? 3 : 2;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int tj_threads = blockDim.x >> 5;
        const int ti = item % p.tiles_i, tj = item / p.tiles_i;
        const int i = ti * 32 + lane, j = tj * tj_threads + warp;
        if (i >= p.ni || j >= p.nj)
            return;
        const int nk = p.nk;
        const T dtr = p.dtr;
        const T bet_m = T(0.5), bet_p = T(0.5); // vertical_advection_defs.hpp
        uint64_t pol_stream = 0, pol_keep = 0;
        if constexpr (Hints) {
            pol_stream = ptx::policy_evict_first();
            pol_keep = ptx::policy_evict_last();
        }

        T *us_p = p.utens_stage.ptr + i + (int64_t)j * p.utens_stage.sj;
        const T *un_p = p.u_stage.ptr + i + (int64_t)j * p.u_stage.sj;
        const T *wc_p = p.wcon.ptr + i + (int64_t)j * p.wcon.sj;
        const T *up_p = p.u_pos.ptr + i + (int64_t)j * p.u_pos.sj;
        const T *ut_p = p.utens.ptr + i + (int64_t)j * p.utens.sj;
        const int64_t us_sk = p.utens_stage.sk, un_sk = p.u_stage.sk, wc_sk = p.wcon.sk, up_sk = p.u_pos.sk,
                      ut_sk = p.utens.sk;

        T *sc;            // ccol(k) at sc[(NS*k) * sc_stride], dcol(k) at sc[(NS*k+1) * sc_stride], u_pos(k) at +2
        int64_t sc_stride;
        if constexpr (SMEM) {
            sc = reinterpret_cast<T *>(smem_raw) + threadIdx.x;
            sc_stride = blockDim.x;
        } else {
            sc = p.scratch + (p.persistent ? (int64_t)blockIdx.x * blockDim.x + threadIdx.x : (int64_t)j * p.ni + i);
            sc_stride = p.slots;
        }
        auto sc_store = [&](int k, T cc, T dc, T up) {
            if constexpr (SMEM || !Hints) {
                sc[(NS * k) * sc_stride] = cc;
                sc[(NS * k + 1) * sc_stride] = dc;
                if constexpr (SAVE_UPOS)
                    sc[(NS * k + 2) * sc_stride] = up;
            } else {
                ptx::st_hint(sc + (NS * k) * sc_stride, cc, pol_keep);
                ptx::st_hint(sc + (NS * k + 1) * sc_stride, dc, pol_keep);
                if constexpr (SAVE_UPOS)
                    ptx::st_hint(sc + (NS * k + 2) * sc_stride, up, pol_keep);
            }
        };
        auto sc_load = [&](int k, T &cc, T &dc, T &up) {
            if constexpr (SMEM || !Hints) {
                cc = sc[(NS * k) * sc_stride];
                dc = sc[(NS * k + 1) * sc_stride];
                if constexpr (SAVE_UPOS)
                    up = sc[(NS * k + 2) * sc_stride];
            } else {
                cc = ptx::ld_hint(sc + (NS * k) * sc_stride, pol_keep);
                dc = ptx::ld_hint(sc + (NS * k + 1) * sc_stride, pol_keep);
                if constexpr (SAVE_UPOS)
                    up = ptx::ld_hint(sc + (NS * k + 2) * sc_stride, pol_keep);
            }
        };

        auto load_level = [&](int k, va_level<T> &v) {
            if (k < nk) {
                v.us = ldg_stream<T, Hints>(us_p + k * us_sk, pol_stream);
                v.up = ldg_stream<T, Hints>(up_p + k * up_sk, pol_stream);
                v.ut = ldg_stream<T, Hints>(ut_p + k * ut_sk, pol_stream);
                if (k + 1 < nk) {
                    v.un = ldg_stream<T, Hints>(un_p + (k + 1) * un_sk, pol_stream);
                    v.w0 = __ldg(wc_p + (k + 1) * wc_sk); // also read by the neighbouring lane: keep it in L1
                    v.w1 = __ldg(wc_p + (k + 1) * wc_sk + 1);
                }
            }
        };

        // ------------------------------------------------------------------ forward sweep (u_forward_function)
        va_level<T> cur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            load_level(u, cur[u]);
        T u_k = ldg_stream<T, Hints>(un_p, pol_stream); // u_stage(k), starts at k = 0
        T u_km1 = T(0);
        T wsum_k = T(0); // wcon(i+1,k) + wcon(i,k)
        T cc_prev = T(0), dc_prev = T(0);
        T up_last = T(0);

        for (int k0 = 0; k0 < nk; k0 += UNROLL) {
            va_level<T> nxt[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                load_level(k0 + UNROLL + u, nxt[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 + u;
                if (k < nk) {
                    const va_level<T> &v = cur[u];
                    T dd = dtr * v.up + v.ut + v.us; // dtr_stage * u_pos + utens + utens_stage
                    T cc, dc;
                    if (k == 0) { // first_level, vertical_advection_dycore.cpp:85-98
                        T wsum_n = v.w1 + v.w0;
                        T gcv = T(.25) * wsum_n;
                        T cs = gcv * bet_m;
                        T c = gcv * bet_p;
                        T b = dtr - c;
                        T correction = -cs * (v.un - u_k);
                        T d = dd + correction;
                        T divided = T(1) / b;
                        cc = c * divided;
                        dc = d * divided;
                        wsum_k = wsum_n;
                    } else if (k < nk - 1) { // body, :50-68
                        T wsum_n = v.w1 + v.w0;
                        T gav = -T(.25) * wsum_k;
                        T gcv = T(.25) * wsum_n;
                        T as = gav * bet_m;
                        T cs = gcv * bet_m;
                        T a = gav * bet_p;
                        T c = gcv * bet_p;
                        T b = dtr - a - c;
                        T correction = -as * (u_km1 - u_k) - cs * (v.un - u_k);
                        T d = dd + correction;
                        T divided = T(1) / (b - cc_prev * a);
                        cc = c * divided;
                        dc = (d - dc_prev * a) * divided;
                        wsum_k = wsum_n;
                    } else { // last_level, :70-83
                        T gav = -T(.25) * wsum_k;
                        T as = gav * bet_m;
                        T a = gav * bet_p;
                        T b = dtr - a;
                        T correction = -as * (u_km1 - u_k);
                        T d = dd + correction;
                        T divided = T(1) / (b - cc_prev * a);
                        cc = cc_prev; // ccol is not written on the last level
                        dc = (d - dc_prev * a) * divided;
                        up_last = v.up;
                    }
                    if (k < nk - 1)
                        sc_store(k, cc, dc, v.up);
                    cc_prev = cc;
                    dc_prev = dc;
                    u_km1 = u_k;
                    u_k = v.un;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                cur[u] = nxt[u];
        }

        // ------------------------------------------------------------------ backward sweep (u_backward_function)
        // last_level :118-121
        T data = dc_prev;
        us_p[(int64_t)(nk - 1) * us_sk] = dtr * (data - up_last);
        struct back_level {
            T cc, dc, up;
        };
        auto load_back = [&](int k, back_level &v) {
            if (k >= 0) {
                sc_load(k, v.cc, v.dc, v.up);
                if constexpr (!SAVE_UPOS)
                    v.up = ldg_stream<T, Hints>(up_p + k * up_sk, pol_stream);
            }
        };
        back_level bcur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            load_back(nk - 2 - u, bcur[u]);
        for (int k0 = nk - 2; k0 >= 0; k0 -= UNROLL) {
            back_level bnxt[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                load_back(k0 - UNROLL - u, bnxt[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                if (k >= 0) { // body :111-116
                    data = bcur[u].dc - bcur[u].cc * data;
                    us_p[(int64_t)k * us_sk] = dtr * (data - bcur[u].up);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                bcur[u] = bnxt[u];
        }
    }

    template <class T, int UNROLL, bool SMEM, bool Hints, bool SAVE_UPOS>
    __global__ void va_kernel(const va_params<T> p) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        if (p.persistent) {
            // resident CTAs walk the column blocks; a thread re-uses its scratch slot for every block it handles,
            // so the scratch footprint is (resident threads) x nk instead of (all columns) x nk and stays in L2
            for (int item = blockIdx.x; item < p.items; item += gridDim.x)
                va_columns<T, UNROLL, SMEM, Hints, SAVE_UPOS>(p, item, smem_raw);
        } else {
            va_columns<T, UNROLL, SMEM, Hints, SAVE_UPOS>(p, blockIdx.x, smem_raw);
        }
    }

    // ------------------------------------------------------------------------------ Thomas solve (tridiagonal.cpp)
    template <class T>
    struct td_params {
        col_field<const T> inf, diag;
        col_field<T> sup, rhs, out;
        int ni, nj, nk;
        int tiles_i;
    };

    template <class T, int UNROLL>
    __global__ void td_kernel(const td_params<T> p) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int tj_threads = blockDim.x >> 5;
        const int ti = blockIdx.x % p.tiles_i, tj = blockIdx.x / p.tiles_i;
        const int i = ti * 32 + lane, j = tj * tj_threads + warp;
        if (i >= p.ni || j >= p.nj)
            return;
        const int nk = p.nk;
        const T *inf_p = p.inf.ptr + i + (int64_t)j * p.inf.sj;
        const T *diag_p = p.diag.ptr + i + (int64_t)j * p.diag.sj;
        T *sup_p = p.sup.ptr + i + (int64_t)j * p.sup.sj;
        T *rhs_p = p.rhs.ptr + i + (int64_t)j * p.rhs.sj;
        T *out_p = p.out.ptr + i + (int64_t)j * p.out.sj;
        struct lvl {
            T inf, diag, sup, rhs;
        };
        auto load = [&](int k, lvl &v) {
            if (k < nk) {
                v.inf = __ldg(inf_p + k * p.inf.sk);
                v.diag = __ldg(diag_p + k * p.diag.sk);
                v.sup = sup_p[k * p.sup.sk];
                v.rhs = rhs_p[k * p.rhs.sk];
            }
        };
        lvl cur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            load(u, cur[u]);
        T sup_prev = T(0), rhs_prev = T(0);
        for (int k0 = 0; k0 < nk; k0 += UNROLL) {
            lvl nxt[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                load(k0 + UNROLL + u, nxt[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 + u;
                if (k < nk) {
                    const lvl &v = cur[u];
                    T s, r;
                    if (k == 0) { // forward_thomas first_level, tridiagonal.cpp:58-62
                        s = v.sup / v.diag;
                        r = v.rhs / v.diag;
                    } else { // :46-56
                        T den = v.diag - sup_prev * v.inf;
                        s = v.sup / den;
                        r = (v.rhs - v.inf * rhs_prev) / den;
                    }
                    sup_p[k * p.sup.sk] = s;
                    rhs_p[k * p.rhs.sk] = r;
                    sup_prev = s;
                    rhs_prev = r;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                cur[u] = nxt[u];
        }
        // backward_thomas :65-74
        T x = rhs_prev;
        out_p[(int64_t)(nk - 1) * p.out.sk] = x;
        struct blvl {
            T sup, rhs;
        };
        auto bload = [&](int k, blvl &v) {
            if (k >= 0) {
                v.sup = sup_p[k * p.sup.sk];
                v.rhs = rhs_p[k * p.rhs.sk];
            }
        };
        blvl bcur[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            bload(nk - 2 - u, bcur[u]);
        for (int k0 = nk - 2; k0 >= 0; k0 -= UNROLL) {
            blvl bnxt[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                bload(k0 - UNROLL - u, bnxt[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                if (k >= 0) {
                    x = bcur[u].rhs - bcur[u].sup * x;
                    out_p[(int64_t)k * p.out.sk] = x;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                bcur[u] = bnxt[u];
        }
    }

    // ------------------------------------------------------------------------------------------ host side
    template <class T, int UNROLL, bool SMEM, bool Hints, bool SAVE_UPOS>
    int launch_va(const va_params<T> &p, int threads, int smem, int grid, cudaStream_t stream) {
        auto kernel = va_kernel<T, UNROLL, SMEM, Hints, SAVE_UPOS>;
        if (smem > 48 * 1024) {
            static thread_local int done_smem = 0, done_dev = -1;
            if (done_smem < smem || done_dev != dev()->device) {
                GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                done_smem = smem;
                done_dev = dev()->device;
            }
        }
        kernel<<<grid, threads, smem, stream>>>(p);
        count_launch();
        return check_launch("va_kernel");
    }

    template <class T, bool SMEM, bool Hints, bool SAVE_UPOS>
    int dispatch_unroll(const va_params<T> &p, int unroll, int threads, int smem, int grid, cudaStream_t stream) {
        switch (unroll) {
        case 1:
            return launch_va<T, 1, SMEM, Hints, SAVE_UPOS>(p, threads, smem, grid, stream);
        case 2:
            return launch_va<T, 2, SMEM, Hints, SAVE_UPOS>(p, threads, smem, grid, stream);
        case 4:
            return launch_va<T, 4, SMEM, Hints, SAVE_UPOS>(p, threads, smem, grid, stream);
        case 8:
            return launch_va<T, 8, SMEM, Hints, SAVE_UPOS>(p, threads, smem, grid, stream);
        default:
            return fail(GTB_ERR_ARG, "gtb_vert_adv: va.unroll must be 1, 2, 4 or 8");
        }
    }

    template <class T, bool SMEM>
    int dispatch_flags(const va_params<T> &p, bool hints, bool save_upos, int unroll, int threads, int smem, int grid,
        cudaStream_t stream) {
        if (hints)
            return save_upos ? dispatch_unroll<T, SMEM, true, true>(p, unroll, threads, smem, grid, stream)
                             : dispatch_unroll<T, SMEM, true, false>(p, unroll, threads, smem, grid, stream);
        return save_upos ? dispatch_unroll<T, SMEM, false, true>(p, unroll, threads, smem, grid, stream)
                         : dispatch_unroll<T, SMEM, false, false>(p, unroll, threads, smem, grid, stream);
    }

    template <class T>
    col_field<T> make_col(const gtb_field *f) {
        return {static_cast<T *>(f->ptr), f->stride_j, f->stride_k};
    }

    template <class T>
    int vert_adv(const gtb_field *utens_stage, const gtb_field *u_stage, const gtb_field *wcon, const gtb_field *u_pos,
        const gtb_field *utens, T dtr_stage, int ni, int nj, int nk, void *stream) {
        const gtb_field *all[5] = {utens_stage, u_stage, wcon, u_pos, utens};
        for (auto f : all) {
            if (!field_ok(f))
                return fail(GTB_ERR_ARG, "gtb_vert_adv: null field");
            if (f->stride_i != 1)
                return fail(GTB_ERR_LAYOUT, "gtb_vert_adv: stride_i must be 1 (i is the unit-stride axis of storage::gpu)");
        }
        if (ni < 0 || nj < 0 || nk < 0)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: negative size");
        if (nk < 2 && ni > 0 && nj > 0)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: nk must be >= 2 (first_level and last_level are distinct levels)");
        device_state *d = dev();
        if (!d)
            return GTB_ERR_CUDA;
        if (ni == 0 || nj == 0)
            return GTB_OK;
        const options &o = opts();
        int threads = o.va_threads ? o.va_threads : 64;
        if (threads % 32 != 0 || threads < 32 || threads > 1024)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: va.threads must be a multiple of 32 in 32..1024");
        int unroll = o.va_unroll ? o.va_unroll : 4;
        va_params<T> p;
        p.utens_stage = make_col<T>(utens_stage);
        p.u_stage = make_col<const T>(u_stage);
        p.wcon = make_col<const T>(wcon);
        p.u_pos = make_col<const T>(u_pos);
        p.utens = make_col<const T>(utens);
        p.dtr = dtr_stage;
        p.ni = ni, p.nj = nj, p.nk = nk;
        p.tiles_i = ceil_div(ni, 32);
        const int64_t items = (int64_t)p.tiles_i * ceil_div(nj, threads / 32);
        if (items >= (int64_t)1 << 31)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: domain too large");
        p.items = (int)items;
        p.scratch = nullptr;
        const bool save_upos = o.va_save_upos != 0;
        const int ns = save_upos ? 3 : 2;
        int grid = p.items;
        p.persistent = 0;
        // va.ctas_per_sm > 0: that many resident CTAs per SM; < 0: an absolute grid size (used by the tests to
        // exercise the persistent path on small domains)
        const int64_t want = o.va_ctas_per_sm > 0 ? (int64_t)o.va_ctas_per_sm * d->sm_count : -(int64_t)o.va_ctas_per_sm;
        if (want > 0 && want < items) {
            grid = (int)want;
            p.persistent = 1;
        }
        const int64_t smem_need = (int64_t)ns * nk * threads * (int64_t)sizeof(T);
        int mode = o.va_scratch;
        if (mode == 0)
            mode = 1;
        if (mode == 2 && smem_need > d->max_smem_optin)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: va.scratch=2 needs %lld bytes of shared memory per CTA (max %d)",
                (long long)smem_need, d->max_smem_optin);
        cudaStream_t s = as_stream(stream);
        if (mode == 2)
            return dispatch_flags<T, true>(p, o.va_hints != 0, save_upos, unroll, threads, (int)smem_need, grid, s);
        p.slots = p.persistent ? (int64_t)grid * threads : (int64_t)ni * nj;
        p.scratch = static_cast<T *>(scratch((size_t)ns * nk * p.slots * sizeof(T)));
        if (!p.scratch)
            return GTB_ERR_ALLOC;
        return dispatch_flags<T, false>(p, o.va_hints != 0, save_upos, unroll, threads, 0, grid, s);
    }

} // namespace

GTB_API int gtb_vert_adv_f64(const gtb_field *utens_stage, const gtb_field *u_stage, const gtb_field *wcon,
    const gtb_field *u_pos, const gtb_field *utens, double dtr_stage, int ni, int nj, int nk, void *stream) {
    return vert_adv<double>(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage, ni, nj, nk, stream);
}

GTB_API int gtb_vert_adv_f32(const gtb_field *utens_stage, const gtb_field *u_stage, const gtb_field *wcon,
    const gtb_field *u_pos, const gtb_field *utens, float dtr_stage, int ni, int nj, int nk, void *stream) {
    return vert_adv<float>(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage, ni, nj, nk, stream);
}

GTB_API int gtb_tridiagonal_f64(const gtb_field *inf, const gtb_field *diag, const gtb_field *sup, const gtb_field *rhs,
    const gtb_field *out, int ni, int nj, int nk, void *stream) {
    const gtb_field *all[5] = {inf, diag, sup, rhs, out};
    for (auto f : all) {
        if (!field_ok(f))
            return fail(GTB_ERR_ARG, "gtb_tridiagonal_f64: null field");
        if (f->stride_i != 1)
            return fail(GTB_ERR_LAYOUT, "gtb_tridiagonal_f64: stride_i must be 1");
    }
    if (ni < 0 || nj < 0 || nk < 0)
        return fail(GTB_ERR_ARG, "gtb_tridiagonal_f64: negative size");
    if (!dev())
        return GTB_ERR_CUDA;
    if (ni == 0 || nj == 0 || nk == 0)
        return GTB_OK;
    td_params<double> p;
    p.inf = make_col<const double>(inf);
    p.diag = make_col<const double>(diag);
    p.sup = make_col<double>(sup);
    p.rhs = make_col<double>(rhs);
    p.out = make_col<double>(out);
    p.ni = ni, p.nj = nj, p.nk = nk;
    p.tiles_i = ceil_div(ni, 32);
    const int threads = 64;
    const int64_t blocks = (int64_t)p.tiles_i * ceil_div(nj, threads / 32);
    td_kernel<double, 4><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(p);
    count_launch();
    return check_launch("td_kernel");
}
