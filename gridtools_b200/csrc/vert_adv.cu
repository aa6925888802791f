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
#include "tma.cuh"

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
        int kc;         // levels per TMA stage (TMA variant)
        int debug;      // diagnosis only (va.debug): 1 skip backward sweep, 2 skip forward math, 4 skip scratch stores
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

    // Rotating k_caches of one column (registers): u_stage(k-1), u_stage(k), wcon(i+1,k)+wcon(i,k), ccol/dcol(k-1).
    template <class T>
    struct va_state {
        T u_k, u_km1, wsum_k, cc_prev, dc_prev, up_last;
    };

    // One level of u_forward_function for one column.  us = utens_stage(k), un = u_stage(k+1), w0/w1 = wcon(i,k+1) /
    // wcon(i+1,k+1), up = u_pos(k), ut = utens(k).  Returns ccol(k), dcol(k) and slides the k_caches.
    template <class T>
    __device__ __forceinline__ void va_forward_level(
        int k, int nk, T dtr, T us, T un, T w0, T w1, T up, T ut, va_state<T> &s, T &cc, T &dc) {
        const T bet_m = T(0.5), bet_p = T(0.5); // vertical_advection_defs.hpp
        T dd = dtr * up + ut + us;                // dtr_stage * u_pos + utens + utens_stage
        if (k == 0) {                             // first_level, vertical_advection_dycore.cpp:85-98
            T wsum_n = w1 + w0;
            T gcv = T(.25) * wsum_n;
            T cs = gcv * bet_m;
            T c = gcv * bet_p;
            T b = dtr - c;
            T correction = -cs * (un - s.u_k);
            T d = dd + correction;
            T divided = T(1) / b;
            cc = c * divided;
            dc = d * divided;
            s.wsum_k = wsum_n;
        } else if (k < nk - 1) { // body, :50-68
            T wsum_n = w1 + w0;
            T gav = -T(.25) * s.wsum_k;
            T gcv = T(.25) * wsum_n;
            T as = gav * bet_m;
            T cs = gcv * bet_m;
            T a = gav * bet_p;
            T c = gcv * bet_p;
            T b = dtr - a - c;
            T correction = -as * (s.u_km1 - s.u_k) - cs * (un - s.u_k);
            T d = dd + correction;
            T divided = T(1) / (b - s.cc_prev * a);
            cc = c * divided;
            dc = (d - s.dc_prev * a) * divided;
            s.wsum_k = wsum_n;
        } else { // last_level, :70-83
            T gav = -T(.25) * s.wsum_k;
            T as = gav * bet_m;
            T a = gav * bet_p;
            T b = dtr - a;
            T correction = -as * (s.u_km1 - s.u_k);
            T d = dd + correction;
            T divided = T(1) / (b - s.cc_prev * a);
            cc = s.cc_prev; // ccol is not written on the last level
            dc = (d - s.dc_prev * a) * divided;
            s.up_last = up;
        }
        s.cc_prev = cc;
        s.dc_prev = dc;
        s.u_km1 = s.u_k;
        s.u_k = un;
    }

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
        va_state<T> st;
        st.u_k = ldg_stream<T, Hints>(un_p, pol_stream); // u_stage(k), starts at k = 0
        st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);

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
                    T cc, dc;
                    va_forward_level<T>(k, nk, dtr, v.us, v.un, v.w0, v.w1, v.up, v.ut, st, cc, dc);
                    if (k < nk - 1)
                        sc_store(k, cc, dc, v.up);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                cur[u] = nxt[u];
        }

        // ------------------------------------------------------------------ backward sweep (u_backward_function)
        // last_level :118-121
        T data = st.dc_prev;
        us_p[(int64_t)(nk - 1) * us_sk] = dtr * (data - st.up_last);
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

    // ------------------------------------------------------------------ TMA-streamed variant (va.variant = 2)
    // One warp = one strip of 32 columns, persistent over the strips.  The warp is its own producer: lane 0 keeps a
    // private ring of S stages in shared memory filled by TMA, each stage holding KC levels of the five fields for
    // the strip ({32,1,KC} boxes; wcon one 16-byte chunk wider for the i+1 neighbour; u_stage and wcon shifted one
    // level up because the functor reads them at k+1).  Nothing but the warp itself touches the ring, so there is no
    // block barrier anywhere; (S-1)*KC levels of HBM traffic stay in flight per warp without using registers, which
    // is what a sequential sweep needs to be bandwidth- instead of latency-bound.  ccol/dcol/u_pos go to a per-warp
    // slab [k][3][32] that is re-used for every strip of the warp: with 7 warps per SM the slabs total ~64 MB and
    // stay in L2 (evict_last), so the flush/fill traffic of the reference's k_caches never reaches HBM.
    template <class T>
    struct va_tma_layout {
        static constexpr int es = (int)sizeof(T);
        static constexpr int ww = 32 + 16 / es; // wcon box width (needs i0 .. i0+32)
        template <int KC>
        static constexpr int stage_bytes() {
            return (KC * (4 * 32 + ww) * es + 127) / 128 * 128;
        }
    };

    struct va_maps {
        CUtensorMap us, up, ut, un, wc;
    };

    template <class T, int KC, int S, int WARPS, int NS>
    __global__ void __launch_bounds__(WARPS * 32) va_tma_kernel(const __grid_constant__ va_maps maps,
        const va_params<T> p) {
        using L = va_tma_layout<T>;
        constexpr int stage_bytes = L::template stage_bytes<KC>();
        extern __shared__ __align__(128) unsigned char smem_all[];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned char *ring = smem_all + warp * (S * stage_bytes);
        uint64_t *full = reinterpret_cast<uint64_t *>(smem_all + WARPS * S * stage_bytes) + warp * S;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s)
                ptx::mbar_init(&full[s], 1);
            ptx::fence_barrier_init();
            if (threadIdx.x == 0) {
                ptx::prefetch_tensormap(&maps.us);
                ptx::prefetch_tensormap(&maps.up);
                ptx::prefetch_tensormap(&maps.ut);
                ptx::prefetch_tensormap(&maps.un);
                ptx::prefetch_tensormap(&maps.wc);
            }
        }
        __syncwarp();
        const int nk = p.nk;
        const T dtr = p.dtr;
        const uint64_t pol_keep = ptx::policy_evict_last();
        const int gw = blockIdx.x * WARPS + warp, total = gridDim.x * WARPS;
        T *slab = p.scratch + (int64_t)gw * 32 + lane; // [k][NS][slots]
        const int64_t sstride = p.slots;
        const int nchunks = (nk + KC - 1) / KC;
        uint32_t n_issued = 0, n_waited = 0; // ring uses so far (stage = n % S, parity = (n / S) & 1)

        for (int item = gw; item < p.items; item += total) {
            const int ti = item % p.tiles_i, j = item / p.tiles_i;
            const int i0 = ti * 32, i = i0 + lane;
            const bool active = i < p.ni;
            // Whole warp: lane 0 arms the barrier, then lanes 0..4 issue the five box loads with ONE predicated
            // instruction (a TMA issue costs ~150 cycles of the issuing thread; five in a row by one lane made the
            // ring refill as expensive as the forward math of the chunk).
            auto issue = [&](int c) {
                const int s = n_issued % S;
                if (lane == 0)
                    ptx::mbar_expect_tx(&full[s], KC * (4 * 32 + L::ww) * L::es);
                __syncwarp();
                if (lane < 5)
                    ptx::tma_load_3d(ring + s * stage_bytes + lane * (KC * 32 * L::es), &maps.us + lane, &full[s], i0, j,
                        c * KC + (lane >= 3 ? 1 : 0)); // u_stage and wcon are read one level up
            };
            // prologue: S-1 chunks in flight
            for (int c = 0; c < S - 1 && c < nchunks; ++c) {
                issue(c);
                ++n_issued;
            }
            va_state<T> st;
            st.u_k = active ? __ldg(p.u_stage.ptr + i + (int64_t)j * p.u_stage.sj) : T(0);
            st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);
            for (int c = 0; c < nchunks; ++c) {
                if (c + S - 1 < nchunks) { // refill the stage consumed in the previous iteration
                    issue(c + S - 1);
                    ++n_issued;
                }
                const int s = n_waited % S;
                ptx::mbar_wait(&full[s], (n_waited / S) & 1);
                ++n_waited;
                const T *sd = reinterpret_cast<const T *>(ring + s * stage_bytes);
                const T *wc = sd + 4 * KC * 32;
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    const int k = c * KC + u;
                    if (k < nk) {
                        T us = sd[u * 32 + lane], up = sd[(KC + u) * 32 + lane], ut = sd[(2 * KC + u) * 32 + lane];
                        T un = sd[(3 * KC + u) * 32 + lane];
                        T w0 = wc[u * L::ww + lane], w1 = wc[u * L::ww + lane + 1];
                        T cc, dc;
                        if (p.debug & 2) {
                            cc = us + un + w0 + w1;
                            dc = up + ut;
                            st.dc_prev = dc;
                        } else {
                            va_forward_level<T>(k, nk, dtr, us, un, w0, w1, up, ut, st, cc, dc);
                        }
                        if (k < nk - 1 && !(p.debug & 4)) {
                            T *q = slab + (int64_t)k * NS * sstride;
                            ptx::st_hint(q, cc, pol_keep);
                            ptx::st_hint(q + sstride, dc, pol_keep);
                            if constexpr (NS == 3)
                                ptx::st_hint(q + 2 * sstride, up, pol_keep);
                        }
                    }
                }
                __syncwarp(); // all lanes are done with stage s before lane 0 refills it
            }
            // ---------------------------------------------------------------- backward sweep (u_backward_function)
            T *us_p = p.utens_stage.ptr + i + (int64_t)j * p.utens_stage.sj;
            const T *up_p = p.u_pos.ptr + (active ? i : 0) + (int64_t)j * p.u_pos.sj;
            const int64_t us_sk = p.utens_stage.sk, up_sk = p.u_pos.sk;
            T data = st.dc_prev;
            if (active)
                us_p[(int64_t)(nk - 1) * us_sk] = dtr * (data - st.up_last);
            constexpr int BU = 8;
            struct back_level {
                T cc, dc, up;
            };
            auto load_back = [&](int k, back_level &v) {
                if (k >= 0) {
                    const T *q = slab + (int64_t)k * NS * sstride;
                    v.cc = ptx::ld_hint(q, pol_keep);
                    v.dc = ptx::ld_hint(q + sstride, pol_keep);
                    if constexpr (NS == 3)
                        v.up = ptx::ld_hint(q + 2 * sstride, pol_keep);
                    else
                        v.up = __ldg(up_p + k * up_sk);
                }
            };
            back_level bcur[BU];
            if (p.debug & 1)
                continue;
#pragma unroll
            for (int u = 0; u < BU; ++u)
                load_back(nk - 2 - u, bcur[u]);
            for (int k0 = nk - 2; k0 >= 0; k0 -= BU) {
                back_level bnxt[BU];
#pragma unroll
                for (int u = 0; u < BU; ++u)
                    load_back(k0 - BU - u, bnxt[u]);
#pragma unroll
                for (int u = 0; u < BU; ++u) {
                    const int k = k0 - u;
                    if (k >= 0) { // body :111-116
                        data = bcur[u].dc - bcur[u].cc * data;
                        if (active)
                            us_p[(int64_t)k * us_sk] = dtr * (data - bcur[u].up);
                    }
                }
#pragma unroll
                for (int u = 0; u < BU; ++u)
                    bcur[u] = bnxt[u];
            }
        }
    }

    // ------------------------------------------------------- forward/backward warp-specialised variant (va.variant = 3)
    // A CTA is a pair of warps working on the same sequence of 32-column strips: warp 0 runs the forward sweeps
    // (TMA ring exactly as above), warp 1 runs the backward sweeps one strip behind.  The k-cache slab is double
    // buffered between them and handed over through two shared-memory mbarriers per buffer (ready: forward done,
    // free: backward done).  With one warp doing both sweeps all warps of the chip alternate in lock-step between an
    // HBM-bound phase (forward) and an L2-bound phase (backward); here the two phases of neighbouring strips overlap,
    // so HBM streams continuously while the backward sweeps drain the slabs out of L2.
    template <class T, int KC, int S, int NS>
    __global__ void __launch_bounds__(64) va_fb_kernel(const __grid_constant__ va_maps maps, const va_params<T> p) {
        using L = va_tma_layout<T>;
        constexpr int stage_bytes = L::template stage_bytes<KC>();
        extern __shared__ __align__(128) unsigned char smem_all[];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned char *ring = smem_all;
        uint64_t *full = reinterpret_cast<uint64_t *>(smem_all + S * stage_bytes);
        uint64_t *ready = full + S; // [2]
        uint64_t *freeb = ready + 2; // [2]
        T *tail = reinterpret_cast<T *>(freeb + 2); // [2 buffers][2 values][32 lanes]: dcol(nk-1), u_pos(nk-1)
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s)
                ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&ready[0], 1);
            ptx::mbar_init(&ready[1], 1);
            ptx::mbar_init(&freeb[0], 1);
            ptx::mbar_init(&freeb[1], 1);
            ptx::fence_barrier_init();
            ptx::prefetch_tensormap(&maps.us);
            ptx::prefetch_tensormap(&maps.up);
            ptx::prefetch_tensormap(&maps.ut);
            ptx::prefetch_tensormap(&maps.un);
            ptx::prefetch_tensormap(&maps.wc);
        }
        __syncthreads();
        const int nk = p.nk;
        const T dtr = p.dtr;
        const uint64_t pol_keep = ptx::policy_evict_last();
        const int64_t sstride = p.slots; // slots = gridDim.x * 2 buffers * 32 lanes
        T *slab0 = p.scratch + ((int64_t)blockIdx.x * 2) * 32 + lane;

        if (warp == 0) {
            // ------------------------------------------------------------ forward warp
            const int nchunks = (nk + KC - 1) / KC;
            uint32_t n_issued = 0, n_waited = 0;
            int n = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                const int b = n & 1;
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i0 = ti * 32, i = i0 + lane;
                const bool active = i < p.ni;
                auto issue = [&](int c) { // whole warp, see va_tma_kernel
                    const int s = n_issued % S;
                    if (lane == 0)
                        ptx::mbar_expect_tx(&full[s], KC * (4 * 32 + L::ww) * L::es);
                    __syncwarp();
                    if (lane < 5)
                        ptx::tma_load_3d(ring + s * stage_bytes + lane * (KC * 32 * L::es), &maps.us + lane, &full[s],
                            i0, j, c * KC + (lane >= 3 ? 1 : 0));
                };
                for (int c = 0; c < S - 1 && c < nchunks; ++c) {
                    issue(c);
                    ++n_issued;
                }
                if (n >= 2) // the backward warp must have drained this buffer (strip n-2)
                    ptx::mbar_wait(&freeb[b], (uint32_t)((n / 2 - 1) & 1));
                T *slab = slab0 + b * 32;
                va_state<T> st;
                st.u_k = active ? __ldg(p.u_stage.ptr + i + (int64_t)j * p.u_stage.sj) : T(0);
                st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);
                for (int c = 0; c < nchunks; ++c) {
                    if (c + S - 1 < nchunks) {
                        issue(c + S - 1);
                        ++n_issued;
                    }
                    const int s = n_waited % S;
                    ptx::mbar_wait(&full[s], (n_waited / S) & 1);
                    ++n_waited;
                    const T *sd = reinterpret_cast<const T *>(ring + s * stage_bytes);
                    const T *wc = sd + 4 * KC * 32;
#pragma unroll
                    for (int u = 0; u < KC; ++u) {
                        const int k = c * KC + u;
                        if (k < nk) {
                            T us = sd[u * 32 + lane], up = sd[(KC + u) * 32 + lane], ut = sd[(2 * KC + u) * 32 + lane];
                            T un = sd[(3 * KC + u) * 32 + lane];
                            T w0 = wc[u * L::ww + lane], w1 = wc[u * L::ww + lane + 1];
                            T cc, dc;
                            va_forward_level<T>(k, nk, dtr, us, un, w0, w1, up, ut, st, cc, dc);
                            if (k < nk - 1) {
                                T *q = slab + (int64_t)k * NS * sstride;
                                ptx::st_hint(q, cc, pol_keep);
                                ptx::st_hint(q + sstride, dc, pol_keep);
                                if constexpr (NS == 3)
                                    ptx::st_hint(q + 2 * sstride, up, pol_keep);
                            }
                        }
                    }
                    __syncwarp();
                }
                tail[(b * 2 + 0) * 32 + lane] = st.dc_prev;
                tail[(b * 2 + 1) * 32 + lane] = st.up_last;
                __threadfence_block();
                __syncwarp();
                if (lane == 0)
                    ptx::mbar_arrive(&ready[b]); // release: slab buffer b and its tail are complete
            }
        } else {
            // ------------------------------------------------------------ backward warp
            int n = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                const int b = n & 1;
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i = ti * 32 + lane;
                const bool active = i < p.ni;
                ptx::mbar_wait(&ready[b], (uint32_t)((n / 2) & 1));
                const T *slab = slab0 + b * 32;
                T *us_p = p.utens_stage.ptr + i + (int64_t)j * p.utens_stage.sj;
                const T *up_p = p.u_pos.ptr + (active ? i : 0) + (int64_t)j * p.u_pos.sj;
                const int64_t us_sk = p.utens_stage.sk, up_sk = p.u_pos.sk;
                T data = tail[(b * 2 + 0) * 32 + lane];
                const T up_last = tail[(b * 2 + 1) * 32 + lane];
                if (active)
                    us_p[(int64_t)(nk - 1) * us_sk] = dtr * (data - up_last);
                constexpr int BU = 8;
                struct back_level {
                    T cc, dc, up;
                };
                auto load_back = [&](int k, back_level &v) {
                    if (k >= 0) {
                        const T *q = slab + (int64_t)k * NS * sstride;
                        v.cc = ptx::ld_cg_hint(q, pol_keep);
                        v.dc = ptx::ld_cg_hint(q + sstride, pol_keep);
                        if constexpr (NS == 3)
                            v.up = ptx::ld_cg_hint(q + 2 * sstride, pol_keep);
                        else
                            v.up = __ldg(up_p + k * up_sk);
                    }
                };
                back_level bcur[BU];
#pragma unroll
                for (int u = 0; u < BU; ++u)
                    load_back(nk - 2 - u, bcur[u]);
                for (int k0 = nk - 2; k0 >= 0; k0 -= BU) {
                    back_level bnxt[BU];
#pragma unroll
                    for (int u = 0; u < BU; ++u)
                        load_back(k0 - BU - u, bnxt[u]);
#pragma unroll
                    for (int u = 0; u < BU; ++u) {
                        const int k = k0 - u;
                        if (k >= 0) { // body :111-116
                            data = bcur[u].dc - bcur[u].cc * data;
                            if (active)
                                us_p[(int64_t)k * us_sk] = dtr * (data - bcur[u].up);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < BU; ++u)
                        bcur[u] = bnxt[u];
                }
                __syncwarp();
                if (lane == 0)
                    ptx::mbar_arrive(&freeb[b]);
            }
        }
    }

    // ------------------------------------------------------- producer/consumer warp pair variant (va.variant = 4)
    // Same ring and slab as va_tma_kernel, but the TMA issue is moved to a second warp of the CTA: issuing a box load
    // blocks the issuing warp for several hundred cycles, which the self-feeding warp of va_tma_kernel pays in series
    // with its recurrence (measured: 9.8 us of 25 us per strip).  The producer warp only waits for a free stage
    // (empty mbarrier, arrived by the consumer after its last read of the stage) and issues; the consumer warp only
    // computes.
    template <class T, int KC, int S, int NS>
    __global__ void __launch_bounds__(64) va_pc_kernel(const __grid_constant__ va_maps maps, const va_params<T> p) {
        using L = va_tma_layout<T>;
        constexpr int stage_bytes = L::template stage_bytes<KC>();
        extern __shared__ __align__(128) unsigned char smem_all[];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned char *ring = smem_all;
        uint64_t *full = reinterpret_cast<uint64_t *>(smem_all + S * stage_bytes);
        uint64_t *empty = full + S;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                ptx::mbar_init(&full[s], 1);
                ptx::mbar_init(&empty[s], 1);
            }
            ptx::fence_barrier_init();
            ptx::prefetch_tensormap(&maps.us);
            ptx::prefetch_tensormap(&maps.up);
            ptx::prefetch_tensormap(&maps.ut);
            ptx::prefetch_tensormap(&maps.un);
            ptx::prefetch_tensormap(&maps.wc);
        }
        __syncthreads();
        const int nk = p.nk;
        const int nchunks = (nk + KC - 1) / KC;
        if (warp == 1) {
            // ------------------------------------------------------------ producer: lanes 0..4 issue one box each
            uint32_t n = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i0 = ti * 32;
                for (int c = 0; c < nchunks; ++c, ++n) {
                    const int s = n % S;
                    if (n >= S)
                        ptx::mbar_wait(&empty[s], (n / S - 1) & 1);
                    if (lane == 0)
                        ptx::mbar_expect_tx(&full[s], KC * (4 * 32 + L::ww) * L::es);
                    __syncwarp();
                    if (lane < 5)
                        ptx::tma_load_3d(ring + s * stage_bytes + lane * (KC * 32 * L::es), &maps.us + lane, &full[s],
                            i0, j, c * KC + (lane >= 3 ? 1 : 0));
                }
            }
            return;
        }
        // ---------------------------------------------------------------- consumer: both sweeps of the strip
        const T dtr = p.dtr;
        const uint64_t pol_keep = ptx::policy_evict_last();
        T *slab = p.scratch + (int64_t)blockIdx.x * 32 + lane; // [k][NS][slots]
        const int64_t sstride = p.slots;
        uint32_t n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int ti = item % p.tiles_i, j = item / p.tiles_i;
            const int i = ti * 32 + lane;
            const bool active = i < p.ni;
            va_state<T> st;
            st.u_k = active ? __ldg(p.u_stage.ptr + i + (int64_t)j * p.u_stage.sj) : T(0);
            st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);
            for (int c = 0; c < nchunks; ++c, ++n) {
                const int s = n % S;
                ptx::mbar_wait(&full[s], (n / S) & 1);
                const T *sd = reinterpret_cast<const T *>(ring + s * stage_bytes);
                const T *wc = sd + 4 * KC * 32;
                T us[KC], up[KC], ut[KC], un[KC], w0[KC], w1[KC];
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    us[u] = sd[u * 32 + lane], up[u] = sd[(KC + u) * 32 + lane], ut[u] = sd[(2 * KC + u) * 32 + lane];
                    un[u] = sd[(3 * KC + u) * 32 + lane];
                    w0[u] = wc[u * L::ww + lane], w1[u] = wc[u * L::ww + lane + 1];
                }
                __syncwarp();
                if (lane == 0)
                    ptx::mbar_arrive(&empty[s]); // the stage is in registers: hand it back before the math
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    const int k = c * KC + u;
                    if (k < nk) {
                        T cc, dc;
                        va_forward_level<T>(k, nk, dtr, us[u], un[u], w0[u], w1[u], up[u], ut[u], st, cc, dc);
                        if (k < nk - 1) {
                            T *q = slab + (int64_t)k * NS * sstride;
                            ptx::st_hint(q, cc, pol_keep);
                            ptx::st_hint(q + sstride, dc, pol_keep);
                            if constexpr (NS == 3)
                                ptx::st_hint(q + 2 * sstride, up[u], pol_keep);
                        }
                    }
                }
            }
            // backward sweep (u_backward_function)
            T *us_p = p.utens_stage.ptr + i + (int64_t)j * p.utens_stage.sj;
            const T *up_p = p.u_pos.ptr + (active ? i : 0) + (int64_t)j * p.u_pos.sj;
            const int64_t us_sk = p.utens_stage.sk, up_sk = p.u_pos.sk;
            T data = st.dc_prev;
            if (active)
                us_p[(int64_t)(nk - 1) * us_sk] = dtr * (data - st.up_last);
            constexpr int BU = 8;
            struct back_level {
                T cc, dc, up;
            };
            auto load_back = [&](int k, back_level &v) {
                if (k >= 0) {
                    const T *q = slab + (int64_t)k * NS * sstride;
                    v.cc = ptx::ld_hint(q, pol_keep);
                    v.dc = ptx::ld_hint(q + sstride, pol_keep);
                    if constexpr (NS == 3)
                        v.up = ptx::ld_hint(q + 2 * sstride, pol_keep);
                    else
                        v.up = __ldg(up_p + k * up_sk);
                }
            };
            back_level bcur[BU];
#pragma unroll
            for (int u = 0; u < BU; ++u)
                load_back(nk - 2 - u, bcur[u]);
            for (int k0 = nk - 2; k0 >= 0; k0 -= BU) {
                back_level bnxt[BU];
#pragma unroll
                for (int u = 0; u < BU; ++u)
                    load_back(k0 - BU - u, bnxt[u]);
#pragma unroll
                for (int u = 0; u < BU; ++u) {
                    const int k = k0 - u;
                    if (k >= 0) { // body :111-116
                        data = bcur[u].dc - bcur[u].cc * data;
                        if (active)
                            us_p[(int64_t)k * us_sk] = dtr * (data - bcur[u].up);
                    }
                }
#pragma unroll
                for (int u = 0; u < BU; ++u)
                    bcur[u] = bnxt[u];
            }
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

    template <class T, int KC, int S, int WARPS, int NS>
    int launch_va_tma(const va_maps &maps, const va_params<T> &p, int grid, cudaStream_t stream) {
        using L = va_tma_layout<T>;
        auto kernel = va_tma_kernel<T, KC, S, WARPS, NS>;
        const int smem = WARPS * (S * L::template stage_bytes<KC>() + S * 8);
        static thread_local int done_dev = -1;
        if (done_dev != dev()->device) {
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done_dev = dev()->device;
        }
        kernel<<<grid, WARPS * 32, smem, stream>>>(maps, p);
        count_launch();
        return check_launch("va_tma_kernel");
    }

    template <class T, int KC, int S, int NS>
    int launch_va_fb(const va_maps &maps, const va_params<T> &p, int grid, cudaStream_t stream) {
        using L = va_tma_layout<T>;
        auto kernel = va_fb_kernel<T, KC, S, NS>;
        const int smem = S * L::template stage_bytes<KC>() + (S + 4) * 8 + 2 * 2 * 32 * (int)sizeof(T);
        static thread_local int done_dev = -1;
        if (done_dev != dev()->device) {
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done_dev = dev()->device;
        }
        kernel<<<grid, 64, smem, stream>>>(maps, p);
        count_launch();
        return check_launch("va_fb_kernel");
    }

    template <class T, int KC, int S, int NS>
    int launch_va_pc(const va_maps &maps, const va_params<T> &p, int grid, cudaStream_t stream) {
        using L = va_tma_layout<T>;
        auto kernel = va_pc_kernel<T, KC, S, NS>;
        const int smem = S * L::template stage_bytes<KC>() + 2 * S * 8;
        static thread_local int done_dev = -1;
        if (done_dev != dev()->device) {
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done_dev = dev()->device;
        }
        kernel<<<grid, 64, smem, stream>>>(maps, p);
        count_launch();
        return check_launch("va_pc_kernel");
    }

    // Builds the five tensor maps; false if any field is not TMA-addressable.
    template <class T>
    bool make_va_maps(va_maps &m, const va_params<T> &p) {
        using L = va_tma_layout<T>;
        const int ni = p.ni, nj = p.nj, nk = p.nk;
        auto one = [&](CUtensorMap *map, const T *ptr, int64_t sj, int64_t sk, int len_i, int box_i, int kc) {
            return make_map<T>(map, ptr, sj, sk, 0, 0, len_i, nj, nk, box_i, 1, kc);
        };
        const int kc = p.kc;
        return one(&m.us, p.utens_stage.ptr, p.utens_stage.sj, p.utens_stage.sk, ni, 32, kc) &&
               one(&m.up, p.u_pos.ptr, p.u_pos.sj, p.u_pos.sk, ni, 32, kc) &&
               one(&m.ut, p.utens.ptr, p.utens.sj, p.utens.sk, ni, 32, kc) &&
               one(&m.un, p.u_stage.ptr, p.u_stage.sj, p.u_stage.sk, ni, 32, kc) &&
               one(&m.wc, p.wcon.ptr, p.wcon.sj, p.wcon.sk, ni + 1, L::ww, kc);
    }

    template <class T>
    int vert_adv_tma(va_params<T> &p, const options &o, device_state *d, cudaStream_t stream, bool *done) {
        *done = false;
        const int kc = o.va_unroll == 8 ? 8 : (o.va_unroll == 2 ? 2 : 4);
        p.kc = kc;
        p.debug = o.va_debug;
        va_maps maps;
        if (!make_va_maps<T>(maps, p))
            return GTB_OK; // not addressable: the caller falls back to the register-prefetch kernel
        const bool fb = o.va_variant == 3; // forward/backward warp pairs
        const int wps = o.va_ctas_per_sm > 0 ? o.va_ctas_per_sm : (fb ? 4 : 7); // (forward) warps per SM
        const int64_t strips = (int64_t)p.tiles_i * p.nj;
        p.items = (int)strips;
        int grid = o.va_ctas_per_sm < 0 ? -o.va_ctas_per_sm : wps * d->sm_count;
        if (grid > strips)
            grid = (int)strips;
        p.slots = (int64_t)grid * 32 * (fb ? 2 : 1);
        const bool save_upos = o.va_save_upos != 2;
        p.scratch = static_cast<T *>(scratch((size_t)(save_upos ? 3 : 2) * p.nk * p.slots * sizeof(T)));
        if (!p.scratch)
            return GTB_ERR_ALLOC;
        p.persistent = 1;
        *done = true;
        {
            const int64_t slab = (int64_t)(save_upos ? 3 : 2) * p.nk * p.slots * (int64_t)sizeof(T);
            int st = set_l2_persist(o.l2_persist_mb < 0 ? slab : (int64_t)o.l2_persist_mb << 20);
            if (st)
                return st;
        }
        if (fb)
            return save_upos ? launch_va_fb<T, 4, 4, 3>(maps, p, grid, stream)
                             : launch_va_fb<T, 4, 4, 2>(maps, p, grid, stream);
        if (o.va_variant == 4) {
            if (kc == 8)
                return save_upos ? launch_va_pc<T, 8, 3, 3>(maps, p, grid, stream)
                                 : launch_va_pc<T, 8, 3, 2>(maps, p, grid, stream);
            return save_upos ? launch_va_pc<T, 4, 4, 3>(maps, p, grid, stream)
                             : launch_va_pc<T, 4, 4, 2>(maps, p, grid, stream);
        }
        if (save_upos) {
            switch (kc) {
            case 2:
                return launch_va_tma<T, 2, 6, 1, 3>(maps, p, grid, stream);
            case 8:
                return launch_va_tma<T, 8, 3, 1, 3>(maps, p, grid, stream);
            default:
                return o.va_stages == 6 ? launch_va_tma<T, 4, 6, 1, 3>(maps, p, grid, stream)
                                        : launch_va_tma<T, 4, 4, 1, 3>(maps, p, grid, stream);
            }
        }
        switch (kc) {
        case 2:
            return launch_va_tma<T, 2, 6, 1, 2>(maps, p, grid, stream);
        case 8:
            return launch_va_tma<T, 8, 3, 1, 2>(maps, p, grid, stream);
        default:
            return launch_va_tma<T, 4, 4, 1, 2>(maps, p, grid, stream);
        }
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
        int unroll = o.va_unroll ? o.va_unroll : 8;
        va_params<T> p;
        p.utens_stage = make_col<T>(utens_stage);
        p.u_stage = make_col<const T>(u_stage);
        p.wcon = make_col<const T>(wcon);
        p.u_pos = make_col<const T>(u_pos);
        p.utens = make_col<const T>(utens);
        p.dtr = dtr_stage;
        p.ni = ni, p.nj = nj, p.nk = nk;
        p.tiles_i = ceil_div(ni, 32);
        if ((int64_t)p.tiles_i * nj >= (int64_t)1 << 31)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: domain too large");
        if (o.va_variant != 1) { // 0 auto / 2: TMA-streamed persistent warps
            bool done = false;
            int st = vert_adv_tma<T>(p, o, d, as_stream(stream), &done);
            if (st || done)
                return st;
            if (o.va_variant >= 2)
                return fail(GTB_ERR_LAYOUT,
                    "gtb_vert_adv: va.variant=2/3 (TMA) needs 16-byte aligned origins and stride_j/stride_k that are "
                    "multiples of 16 bytes");
        }
        const int64_t items = (int64_t)p.tiles_i * ceil_div(nj, threads / 32);
        if (items >= (int64_t)1 << 31)
            return fail(GTB_ERR_ARG, "gtb_vert_adv: domain too large");
        p.items = (int)items;
        p.scratch = nullptr;
        const bool save_upos = o.va_save_upos != 2; // 0 auto (on), 1 on, 2 off
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
