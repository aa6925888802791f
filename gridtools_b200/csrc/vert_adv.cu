// vert_adv.cu -- vertical sweeps for sm_100a: vertical_advection_dycore (vertical_advection_dycore.cpp:32-149)
// and the plain Thomas solve (tridiagonal.cpp:39-97).
//
// What the reference does (stencil/gpu/make_kernel_fun.hpp:50-98, k_cache.hpp:24-46, fill_flush.hpp:123-324): one
// thread per column, forward and backward MSS fused into one launch, u_stage in a 3-deep register k_cache filled
// level by level, ccol/dcol in 2-deep register k_caches that are flushed to a blocked global temporary and read back
// by the backward sweep; no prefetching ("no unrolling", make_kernel_fun.hpp:62-64), so every level waits for its
// own loads.
//
// What this file does instead ("va.variant" option; 0 = auto; measured numbers in profiles/README.md):
//  7  va_pair_kernel     fp64 default: one 16-warp CTA per SM = 8 forward / backward warp pairs, ccol/dcol in TENSOR
//                        MEMORY (one window per pair, levels past it in a shared-memory slab), forward inputs through a
//                        per-warp TMA ring, backward sweep of strip n beside the forward sweep of strip n+1;
//  3  va_stream_kernel   fp32 default and columns too tall for 7: persistent one-warp CTAs, per-warp TMA ring,
//                        ccol/dcol/u_pos in a contiguous per-warp L2 slab, 32-level register ring for the backward sweep;
//  1  va_tma_kernel      first TMA version (per-thread slab loads); kept as the reference point of the sweeps;
//  0/- va_kernel         any alignment: one thread per column, loads software-pipelined UNROLL levels ahead through
//                        registers (plain LDG), ccol/dcol in shared memory or a column-interleaved global scratch.
// (Variants 2, 4, 5, 6 of round 1 -- measured slower -- were removed; their numbers are in profiles/r01_*.)
// Common to all: a warp covers 32 consecutive i (256 B per fp64 row), the k_caches of the reference are registers
// rotated by the sweep (u_stage(k-1,k,k+1), wcon(i,k)+wcon(i+1,k), ccol/dcol(k-1), data_col(k+1)), and the
// temporaries ccol/dcol that the reference flushes to HBM stay on the SM (7) or in L2 (3, evict_last).
//
// Arithmetic follows the functor bodies operation by operation (-fmad=false, IEEE division), so results are
// bit-identical to oracle/gt_oracle.c compiled with -ffp-contract=off.
#include "common.cuh"
#include "tma.cuh"

#include <mutex>

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
        int k_split;    // resident variant: levels [0, k_split) keep ccol/dcol in shared memory, the rest in registers
        int debug;      // diagnosis only (va.debug): 1 skip backward sweep, 2 skip forward math, 4 skip scratch stores
        int stages;     // TMEM variants: TMA ring depth (run-time)
        int bstages;    // fused TMEM variant: depth of the u_pos ring of the backward sweep
        int pairs;      // paired-warp variant: F/B warp pairs in use per CTA
        int stagger_ns; // TMEM variant: warp w of a CTA starts w * stagger_ns late (de-phases forward and backward sweeps)
        int *tickets;   // TMEM variants: {next strip, finished warps}, self-resetting; one of 16 slots per launch
        long long *trace; // diagnosis only (va.debug & 128): time stamps per warp pair
        stencil_gate gate; // device-side ordering against a concurrent halo exchange (paired-warp kernel only)
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

    // KC consecutive BODY levels (:50-68) in two phases: first everything that does not depend on the previous level's
    // ccol/dcol (the tridiagonal coefficients a, b, c and the right-hand side d of all KC levels: independent
    // instruction streams the scheduler can interleave), then the recurrence itself, which is the only serial part
    // (multiply, subtract, reciprocal, multiply per level).  Same operations on the same operands as
    // va_forward_level, hence bit-identical results; only the order of independent instructions differs.
    // lv(u, us, un, w0, w1, up, ut) loads level u of the chunk, out(u, cc, dc, up) stores its results.
    template <class T, int KC, class Load, class Store>
    __device__ __forceinline__ void va_forward_body_chunk(T dtr, va_state<T> &s, Load &&lv, Store &&out) {
        const T bet_m = T(0.5), bet_p = T(0.5); // vertical_advection_defs.hpp
        T a[KC], b[KC], c[KC], d[KC], upv[KC];
        T wsum = s.wsum_k, u_km1 = s.u_km1, u_k = s.u_k;
#pragma unroll
        for (int u = 0; u < KC; ++u) {
            T us, un, w0, w1, up, ut;
            lv(u, us, un, w0, w1, up, ut);
            T dd = dtr * up + ut + us;
            T wsum_n = w1 + w0;
            T gav = -T(.25) * wsum;
            T gcv = T(.25) * wsum_n;
            T as = gav * bet_m;
            T cs = gcv * bet_m;
            a[u] = gav * bet_p;
            c[u] = gcv * bet_p;
            b[u] = dtr - a[u] - c[u];
            T correction = -as * (u_km1 - u_k) - cs * (un - u_k);
            d[u] = dd + correction;
            upv[u] = up;
            wsum = wsum_n;
            u_km1 = u_k;
            u_k = un;
        }
        T cc_prev = s.cc_prev, dc_prev = s.dc_prev;
#pragma unroll
        for (int u = 0; u < KC; ++u) {
            T divided = T(1) / (b[u] - cc_prev * a[u]);
            T cc = c[u] * divided;
            T dc = (d[u] - dc_prev * a[u]) * divided;
            out(u, cc, dc, upv[u]);
            cc_prev = cc;
            dc_prev = dc;
        }
        s.cc_prev = cc_prev;
        s.dc_prev = dc_prev;
        s.wsum_k = wsum;
        s.u_km1 = u_km1;
        s.u_k = u_k;
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

    // ------------------------------------------------------------------ streaming variant (va.variant = 3, default)
    // Same decomposition as variant 2 (one warp = one 32-column strip, persistent over the strips, the warp is its
    // own TMA producer), rebuilt around what the profile of variant 2 showed (profiles/README.md):
    //  * 152 warp instructions per level of which 36 are fp64 -- the rest was 64-bit address arithmetic, the
    //    three-way first/body/last branch in every level and register shuffling -- so a warp advanced one level every
    //    ~700 cycles.  Here the k-cache slab of a warp is CONTIGUOUS, [level][ccol, dcol, u_pos][32 lanes] (all stores
    //    of a level are one running pointer plus compile-time offsets), chunks that hold neither the first nor the
    //    last level run a branch-free body, and one lane issues the five TMA boxes of a stage back to back;
    //  * the backward sweep took 20 of 71 us although it is 4 flops per level: all warps of the chip run forward and
    //    backward in lock-step, and a backward sweep that prefetches 8 levels ahead makes 10 dependent L2 round trips
    //    of ~1 us while HBM idles.  With 7 warps per SM a thread may use ~290 registers, so the backward sweep here
    //    keeps NB*8 levels (24 KB per warp at NB = 4) in flight in a statically indexed register ring;
    //  * the forward ring of the NEXT strip is primed before the backward sweep of the current one starts, so HBM
    //    streams while the warp drains its slab out of L2.
    // NS = 3: u_pos(k) travels through the slab next to ccol/dcol (no second HBM read); NS = 2: re-read from HBM.
    template <class T, int KC, int S>
    constexpr int va_stream_smem() {
        return S * va_tma_layout<T>::template stage_bytes<KC>() + S * 8;
    }

    template <class T, int KC, int S, int NB, int NS>
    __global__ void __launch_bounds__(32) va_stream_kernel(const __grid_constant__ va_maps maps, const va_params<T> p) {
        using L = va_tma_layout<T>;
        constexpr int es = L::es;
        constexpr int fstage = L::template stage_bytes<KC>();
        constexpr uint32_t ftx = KC * (4 * 32 + L::ww) * es;
        constexpr int G = 8;       // levels per register block of the backward sweep
        constexpr int D = NB * G;  // levels in flight
        extern __shared__ __align__(128) unsigned char smem_all[];
        unsigned char *fring = smem_all;
        uint64_t *ffull = reinterpret_cast<uint64_t *>(smem_all + S * fstage);
        const int lane = threadIdx.x;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s)
                ptx::mbar_init(&ffull[s], 1);
            ptx::fence_barrier_init();
            ptx::prefetch_tensormap(&maps.us);
            ptx::prefetch_tensormap(&maps.up);
            ptx::prefetch_tensormap(&maps.ut);
            ptx::prefetch_tensormap(&maps.un);
            ptx::prefetch_tensormap(&maps.wc);
        }
        __syncwarp();
        const int nk = p.nk;
        const T dtr = p.dtr;
        const int dbg = p.debug; // diagnosis only (va.debug): 1 skip backward, 2 skip forward math, 4 skip slab stores, 8 skip output stores
        const uint64_t pol_keep = ptx::policy_evict_last();
        const int gw = blockIdx.x, total = gridDim.x;
        T *const slab = p.scratch + (int64_t)gw * p.slots + lane; // p.slots = elements per warp slab: nk*NS*32
        const int nchunks = (nk + KC - 1) / KC;
        const int c_tail = (nk - 1) / KC; // first forward chunk that needs per-level checks (holds level nk-1)
        int f_issue = 0, f_wait = 0; // ring stage the producer fills next / the consumer waits for next
        uint32_t f_phase = 0;        // parity of the consumer's current trip around the ring

        // One lane issues the five boxes of a forward stage back to back (UTMALDG takes uniform operands: letting
        // five lanes issue one box each only makes the compiler serialise them in an election loop).
        auto issue_f = [&](int i0, int j, int c) {
            const int s = f_issue;
            f_issue = f_issue + 1 == S ? 0 : f_issue + 1;
            if (ptx::elect_one()) {
                unsigned char *st = fring + s * fstage;
                uint64_t *bar = &ffull[s];
                ptx::mbar_expect_tx(bar, ftx);
                ptx::tma_load_3d(st, &maps.us, bar, i0, j, c * KC);
                ptx::tma_load_3d(st + KC * 32 * es, &maps.up, bar, i0, j, c * KC);
                ptx::tma_load_3d(st + 2 * KC * 32 * es, &maps.ut, bar, i0, j, c * KC);
                ptx::tma_load_3d(st + 3 * KC * 32 * es, &maps.un, bar, i0, j, c * KC + 1); // read one level up
                ptx::tma_load_3d(st + 4 * KC * 32 * es, &maps.wc, bar, i0, j, c * KC + 1);
            }
        };
        auto prime = [&](int item, T &u0) { // first S-1 forward chunks of a strip + u_stage(k = 0)
            const int ti = item % p.tiles_i, j = item / p.tiles_i;
            const int i0 = ti * 32;
            for (int c = 0; c < S - 1 && c < nchunks; ++c)
                issue_f(i0, j, c);
            u0 = i0 + lane < p.ni ? __ldg(p.u_stage.ptr + i0 + lane + (int64_t)j * p.u_stage.sj) : T(0);
        };

        int item = gw;
        T u0 = T(0);
        if (item < p.items)
            prime(item, u0);
        for (; item < p.items; item += total) {
            const int ti = item % p.tiles_i, j = item / p.tiles_i;
            const int i0 = ti * 32, i = i0 + lane;
            const bool active = i < p.ni;
            va_state<T> st;
            st.u_k = u0;
            st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);
            T *q = slab;
            // ---------------------------------------------------------------- forward sweep (u_forward_function)
            for (int c = 0; c < nchunks; ++c) {
                if (c + S - 1 < nchunks) // refill the stage consumed in the previous iteration
                    issue_f(i0, j, c + S - 1);
                const int s = f_wait;
                ptx::mbar_wait(&ffull[s], f_phase);
                if (++f_wait == S) {
                    f_wait = 0;
                    f_phase ^= 1;
                }
                const T *sd = reinterpret_cast<const T *>(fring + s * fstage);
                const T *wc = sd + 4 * KC * 32;
                if (c != 0 && c < c_tail) {
                    va_forward_body_chunk<T, KC>(
                        dtr, st,
                        [&](int u, T &us, T &un, T &w0, T &w1, T &up, T &ut) {
                            us = sd[u * 32 + lane], up = sd[(KC + u) * 32 + lane], ut = sd[(2 * KC + u) * 32 + lane];
                            un = sd[(3 * KC + u) * 32 + lane];
                            w0 = wc[u * L::ww + lane], w1 = wc[u * L::ww + lane + 1];
                        },
                        [&](int u, T cc, T dc, T up) {
                            if (!(dbg & 4)) {
                                ptx::st_hint(q + (u * NS) * 32, cc, pol_keep);
                                ptx::st_hint(q + (u * NS + 1) * 32, dc, pol_keep);
                                if constexpr (NS == 3)
                                    ptx::st_hint(q + (u * NS + 2) * 32, up, pol_keep);
                            }
                        });
                } else {
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
                                ptx::st_hint(q + (u * NS) * 32, cc, pol_keep);
                                ptx::st_hint(q + (u * NS + 1) * 32, dc, pol_keep);
                                if constexpr (NS == 3)
                                    ptx::st_hint(q + (u * NS + 2) * 32, up, pol_keep);
                            }
                        }
                    }
                }
                q += KC * NS * 32;
                __syncwarp(); // all lanes are done with stage s before lane 0 refills it
            }
            // ---------------------------------------------------------------- backward sweep (u_backward_function)
            // Register ring: block b holds levels kb - b*G - g (g = 0..G-1) of the current window of D levels.
            if (dbg & 1) {
                if (item + total < p.items)
                    prime(item + total, u0);
                if (active)
                    p.utens_stage.ptr[i + (int64_t)j * p.utens_stage.sj] = st.dc_prev;
                continue;
            }
            T bc[NB][G], bd[NB][G], bu[NB][G];
            const int64_t us_sk = p.utens_stage.sk, up_sk = p.u_pos.sk;
            T *o = p.utens_stage.ptr + i + (int64_t)j * p.utens_stage.sj + (int64_t)(nk - 1) * us_sk;
            const T *upp = p.u_pos.ptr + (active ? i : 0) + (int64_t)j * p.u_pos.sj; // NS == 2 only
            auto load_block = [&](int b, int ktop) { // levels ktop .. ktop-G+1
                const T *r = slab + (int64_t)ktop * (NS * 32);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (ktop - g >= 0) {
                        bc[b][g] = ptx::ld_hint(r - g * (NS * 32), pol_keep);
                        bd[b][g] = ptx::ld_hint(r - g * (NS * 32) + 32, pol_keep);
                        if constexpr (NS == 3)
                            bu[b][g] = ptx::ld_hint(r - g * (NS * 32) + 64, pol_keep);
                        else
                            bu[b][g] = __ldg(upp + (int64_t)(ktop - g) * up_sk);
                    }
                }
            };
#pragma unroll
            for (int b = 0; b < NB; ++b)
                load_block(b, nk - 2 - b * G);
            if (item + total < p.items) // HBM keeps streaming while this warp drains its slab
                prime(item + total, u0);
            T data = st.dc_prev; // last_level :118-121
            if (active)
                *o = dtr * (data - st.up_last);
            for (int kb = nk - 2; kb >= 0; kb -= D) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int ktop = kb - b * G;
#pragma unroll
                    for (int g = 0; g < G; ++g) { // body :111-116
                        if (ktop - g >= 0) {
                            data = bd[b][g] - bc[b][g] * data;
                            o -= us_sk;
                            if (active && !(dbg & 8))
                                *o = dtr * (data - bu[b][g]);
                        }
                    }
                    load_block(b, ktop - D);
                }
            }
        }
    }

    // bytes of one stage of the backward sweep's u_pos ring (paired-warp kernel)
    template <class T, int KC>
    constexpr int va_bstage_bytes() {
        return (KC * 32 * (int)sizeof(T) + 127) / 128 * 128;
    }

    // ------------------------------------------------------------------ TMEM variant (va.variant = 5)
    // Tensor memory as the k-cache store.  The profile of variant 3 (profiles/README.md) shows the sweep bound by L2
    // slice traffic: the ccol/dcol/u_pos slab adds 2 x 24 B per point of L2 writes and reads to the 48 B that come
    // from HBM.  Variant 4 keeps ccol/dcol on the SM but pays for it with statically indexed register tiers
    // (jump tables, 121 instructions per level).  An SM of this chip has a third large on-chip memory that a stencil
    // otherwise never touches: 256 KB of tensor memory, 128 lanes x 512 32-bit columns, addressed with a RUNTIME
    // column index by tcgen05.st / tcgen05.ld (32x32b shape: lane l of warp w owns TMEM lane 32*(w%4)+l).  One
    // 32-bit column per lane and word is exactly the shape of a per-column k-cache:
    //  * one CTA per SM with WARPS (7 or 8) independent warps, each the streaming warp of variant 3 (own TMA ring,
    //    own strips, no block barrier in the sweep); warp w owns lane quarter w%4 and column half w/4 of the CTA's
    //    512-column allocation: 256 columns = 64 fp64 levels (128 fp32 levels) of {ccol, dcol};
    //  * a forward chunk of KC levels ends with ONE tcgen05.st.x16 (fp64, KC = 4), a backward chunk starts with ONE
    //    tcgen05.ld.x16 + wait::ld -- no address arithmetic, no L2 traffic, no shared-memory bandwidth;
    //  * levels above the TMEM capacity (64..nk-2 for fp64) go to a small per-warp shared-memory slab;
    //  * u_pos(k) for the backward sweep is re-read with an evict_first load from L2, where the forward TMA load
    //    left it with an evict_last hint; these loads run 32 levels ahead in a register ring.
    // L2 traffic per point: 48 B from HBM + 8 B u_pos re-read + 8 B written, against 96 + 48 in variant 3.
    namespace tm {
        __device__ __forceinline__ void alloc(uint32_t *slot, uint32_t cols) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_addr(slot)),
                         "r"(cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __device__ __forceinline__ void dealloc(uint32_t addr, uint32_t cols) {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
        }
        __device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
        __device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        __device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
        __device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
        template <int N>
        __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[N]) {
            static_assert(N == 8 || N == 16 || N == 32, "tcgen05.st width");
            if constexpr (N == 8)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a),
                             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                             : "memory");
            else if constexpr (N == 16)
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
                    "%13, %14, %15, %16};" ::"r"(a),
                    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                    : "memory");
            else
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
                    "%13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32};" ::"r"(a),
                    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
                    "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
                    "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                    : "memory");
        }
        template <int N>
        __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[N]) {
            static_assert(N == 8 || N == 16 || N == 32, "tcgen05.ld width");
            if constexpr (N == 8)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                             "=r"(r[7])
                             : "r"(a)
                             : "memory");
            else if constexpr (N == 16)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                    "%14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(a)
                    : "memory");
            else
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                    "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                    "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                    "=r"(r[30]), "=r"(r[31])
                    : "r"(a)
                    : "memory");
        }
        // {ccol, dcol} of one level <-> CPL 32-bit words
        template <class T>
        __device__ __forceinline__ void pack(T cc, T dc, uint32_t *w) {
            if constexpr (sizeof(T) == 8) {
                w[0] = (uint32_t)__double2loint(cc), w[1] = (uint32_t)__double2hiint(cc);
                w[2] = (uint32_t)__double2loint(dc), w[3] = (uint32_t)__double2hiint(dc);
            } else {
                w[0] = __float_as_uint(cc), w[1] = __float_as_uint(dc);
            }
        }
        template <class T>
        __device__ __forceinline__ void unpack(const uint32_t *w, T &cc, T &dc) {
            if constexpr (sizeof(T) == 8) {
                cc = __hiloint2double((int)w[1], (int)w[0]);
                dc = __hiloint2double((int)w[3], (int)w[2]);
            } else {
                cc = __uint_as_float(w[0]), dc = __uint_as_float(w[1]);
            }
        }
    } // namespace tm

    template <class T>
    struct va_tmem_cfg {
        static constexpr int cpl = 2 * (int)sizeof(T) / 4; // 32-bit columns per level
        static constexpr int cols(int warps) { return warps <= 4 ? 512 : 256; } // columns per warp
        static constexpr int levels(int warps) { return cols(warps) / cpl; }     // levels a warp keeps in TMEM
    };

    // shared memory: [warps][stages][fstage] | [warps][slab_bytes] | [warps][stages] mbarriers | TMEM base address
    __device__ __forceinline__ uint64_t globaltimer_ns() {
        uint64_t t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }

    // Strips are handed out dynamically: the first one is the warp's own number, the following ones come from a
    // ticket counter (one atomic per strip), so warps that start late (stagger) or run on a slower SM take fewer.
    // The last warp to finish resets the counters for the next launch on the stream.
    // ------------------------------------------------------------------ paired-warp TMEM variant (va.variant = 7)
    // What the debug knobs of variant 5 show (tools/va_dbg.py, profiles/README.md): with ALL arithmetic removed the
    // kernel still takes 58 of its 60 us -- 37 us for streaming the forward inputs, 9 us for the backward sweeps'
    // u_pos re-reads and 11 us for their output stores.  The sweeps are bound by the order in which one warp walks
    // through its memory operations, not by fp64 latency: while a warp sweeps back it streams nothing, and variant 6
    // (both sweeps in one instruction stream) only trades that bubble for a 25 % longer forward chunk.  The SM,
    // meanwhile, issues on 30 % of its cycles and has room for 64 warps.
    // Here every strip slot is a PAIR of warps that share one k-cache window:
    //  * the F warp (warps 0..7) runs forward sweeps back to back: its TMA ring never drains, at the end of a strip
    //    it hands {strip, dcol(nk-1), u_pos(nk-1)} to its partner through shared memory and a named barrier and starts
    //    the next strip at once;
    //  * the B warp (warps 8..15: same TMEM lane quarter as its partner) sweeps the previous strip back WHILE the
    //    F warp fills the window with the next one -- the alternating natural/mirrored chunk order of variant 6 makes
    //    the slot the backward sweep frees the slot the forward sweep fills next; a monotonic chunk counter in
    //    shared memory keeps the F warp from overwriting a slot that has not been read (the B warp is ~4x faster per
    //    chunk, so this never blocks in practice);
    //  * u_pos(k) reaches the B warp through its own small TMA ring (L2 hits, evict_first).
    // Output stores, L2 re-reads and HBM streaming now overlap for the whole launch; only the very first forward
    // sweep and the very last backward sweep of a pair run alone.
    template <class T, int KC>
    int va_pair_smem(int pairs, int stages, int bstages, int slab_bytes) {
        return pairs * (stages * va_tma_layout<T>::template stage_bytes<KC>() + bstages * va_bstage_bytes<T, KC>() +
                           slab_bytes + 4 * 32 * (int)sizeof(T) + (stages + bstages) * 8 + 32) + 16;
    }

    // Named barrier between the two warps of a pair (the non-.aligned form: the warps reach it from different code).
    __device__ __forceinline__ void named_barrier_sync(int id, int threads) {
        __syncwarp();
        asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
    }

    // Backward sweep of one strip with u_pos read straight into registers (va_pair_kernel<..., BLDG = true>): see there.
    template <int N>
    struct va_words {
        uint32_t v[N];
    };
    template <class T, int N>
    struct va_vals {
        T v[N];
    };
    template <class T, int KC>
    __device__ __forceinline__ void va_pair_backward_ldg(const va_params<T> &p, int lane, int i0, int j, int cb_top,
        int c_split, int mirrored, int more, uint32_t tw, T *ss, volatile int *ctl, T data, T up_last, int &done,
        uint64_t pol_stream) {
        using CFG = va_tmem_cfg<T>;
        constexpr int CPL = CFG::cpl;
        constexpr int NW = KC * CPL;
        const int nk = p.nk;
        const T dtr = p.dtr;
        const bool active = i0 + lane < p.ni;
        const int64_t us_sk = p.utens_stage.sk;
        // `upq` walks down the column one chunk per load; only the top chunk can hold levels past nk-2 (never used:
        // clamped to its first level).  (Walking level by level with one pointer, or loads without the L2 hint, are
        // slower: profiles/r02_va_timeline.txt.)
        const int64_t up_sk = p.u_pos.sk;
        const T *upq = p.u_pos.ptr + (active ? i0 + lane : 0) + (int64_t)j * p.u_pos.sj + (int64_t)(cb_top * KC) * up_sk;
        int up_left = cb_top + 1; // chunks not loaded yet
        auto load_up = [&](va_vals<T, KC> &up, auto top) {
            if (up_left <= 0)
                return;
#pragma unroll
            for (int u = 0; u < KC; ++u) {
                const bool in = !decltype(top)::value || cb_top * KC + u <= nk - 2;
                up.v[u] = ptx::ld_hint(in ? upq + u * up_sk : upq, pol_stream);
            }
            upq -= KC * up_sk;
            --up_left;
        };
        va_vals<T, KC> up0, up1, up2, up3;
        load_up(up0, std::true_type());
        load_up(up1, std::false_type());
        load_up(up2, std::false_type());
        load_up(up3, std::false_type());
        T *o = p.utens_stage.ptr + (active ? i0 + lane : 0) + (int64_t)j * p.utens_stage.sj + (int64_t)(nk - 1) * us_sk;
        if (active) // last_level :118-121
            *o = dtr * (data - up_last);
        // {ccol, dcol} of a chunk live in one of two register sets (packed words, as they come out of TMEM): the set of
        // chunk cb-1 is filled while the levels of chunk cb are computed from the other one.
        va_words<NW> w0, w1;
        auto fetch = [&](int cb, va_words<NW> &w) { // issue the loads of chunk cb
            const int ph = mirrored ? cb_top - cb : cb; // where the strip stored its chunk cb
            if (ph < c_split)
                tm::ld<NW>(tw + (uint32_t)(ph * NW), w.v);
            return ph;
        };
        auto land = [&](int cb, int ph, va_words<NW> &w) { // ... and wait for them
            if (ph < c_split) {
                tm::wait_ld();
            } else {
                const T *q = ss + (ph - c_split) * (KC * 64);
#pragma unroll
                for (int u = 0; u < KC; ++u) {
                    T cc = T(0), dc = T(0);
                    if (cb * KC + u < nk - 1) {
                        cc = q[u * 64];
                        dc = q[u * 64 + 32];
                    }
                    tm::pack<T>(cc, dc, w.v + u * CPL);
                }
            }
            ++done;
            if (more && ((done & 1) == 0 || cb == 0)) { // the slots read so far may be overwritten by the partner
                tm::fence_before();
                __threadfence_block();
                __syncwarp();
                if (lane == 0)
                    ctl[0] = done;
            }
        };
        auto sweep = [&](int cb, va_words<NW> &wc, va_words<NW> &wn, va_vals<T, KC> &up, auto checked) {
            int ph_next = 0;
            if (cb > 0)
                ph_next = fetch(cb - 1, wn);
#pragma unroll
            for (int u = KC - 1; u >= 0; --u) { // levels of chunk cb, top down (body :111-116)
                if (!decltype(checked)::value || cb * KC + u <= nk - 2) {
                    T cc, dc;
                    tm::unpack<T>(wc.v + u * CPL, cc, dc);
                    data = dc - cc * data;
                    o -= us_sk;
                    if (active)
                        *o = dtr * (data - up.v[u]);
                }
            }
            load_up(up, std::false_type()); // this register set again four chunks further down
            if (cb > 0)
                land(cb - 1, ph_next, wn);
        };
        land(cb_top, fetch(cb_top, w0), w0);
        sweep(cb_top, w0, w1, up0, std::true_type()); // the top chunk may hold fewer than KC levels
        for (int cb = cb_top - 1; cb >= 0;) {
            sweep(cb--, w1, w0, up1, std::false_type());
            if (cb < 0)
                break;
            sweep(cb--, w0, w1, up2, std::false_type());
            if (cb < 0)
                break;
            sweep(cb--, w1, w0, up3, std::false_type());
            if (cb < 0)
                break;
            sweep(cb--, w0, w1, up0, std::false_type());
        }
    }

    // BLDG: the B warp reads u_pos(k) with plain (L2-hinted) loads into registers, three chunks ahead, instead of through
    // its own TMA ring: per chunk 4 LDGs replace the elected-lane TMA issue (28 instructions), the mbarrier wait and the
    // shared-memory reads -- the backward sweep is bound by the length of its own instruction stream (ncu: `wait` and
    // `branch_resolving` stalls, 160 instructions per 4-level chunk at one warp per scheduler), not by memory.
    template <class T, int KC, bool BLDG>
    __global__ void __launch_bounds__(512, 1) va_pair_kernel(const __grid_constant__ va_maps maps, const va_params<T> p) {
        using L = va_tma_layout<T>;
        using CFG = va_tmem_cfg<T>;
        constexpr int es = L::es;
        constexpr int fstage = L::template stage_bytes<KC>();
        constexpr int bstage = va_bstage_bytes<T, KC>();
        constexpr uint32_t ftx = KC * (4 * 32 + L::ww) * es;
        constexpr uint32_t btx = KC * 32 * es;
        constexpr int CPL = CFG::cpl;
        constexpr int NW = KC * CPL; // 32-bit words of a chunk
        extern __shared__ __align__(128) unsigned char smem_all[];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, PAIRS = p.pairs; // 16 warps, PAIRS <= 8 pairs in use
        const int pw = warp & 7;             // pair index in the CTA: warps pw and pw + 8 share a TMEM lane quarter
        const bool is_b = warp >= 8;         // role
        const bool idle = pw >= PAIRS;
        const int S = p.stages, SB = p.bstages;
        const int slab_bytes = (int)p.slots; // per pair, multiple of 128
        const int per_pair = S * fstage + SB * bstage + slab_bytes + 4 * 32 * es;
        unsigned char *fring = smem_all + pw * per_pair;
        unsigned char *bring = fring + S * fstage;
        T *const ss = reinterpret_cast<T *>(bring + SB * bstage) + lane;               // [chunk][level][2][32]
        T *const hand = reinterpret_cast<T *>(bring + SB * bstage + slab_bytes) + lane; // [parity][dcol, u_pos][32]
        uint64_t *ffull = reinterpret_cast<uint64_t *>(smem_all + PAIRS * per_pair) + pw * (S + SB);
        uint64_t *bfull = ffull + S;
        volatile int *ctl = reinterpret_cast<volatile int *>(smem_all + PAIRS * (per_pair + (S + SB) * 8)) + pw * 8;
        // ctl[0] chunks the B warp has read so far (monotonic), ctl[2 + 3*parity ..]: {strip, more, mirrored}
        uint32_t *tslot = reinterpret_cast<uint32_t *>(smem_all + PAIRS * (per_pair + (S + SB) * 8 + 32));
        if (lane == 0 && !is_b && !idle) {
            for (int s = 0; s < S + SB; ++s)
                ptx::mbar_init(&ffull[s], 1);
            ctl[0] = 0;
            ptx::fence_barrier_init();
            if (warp == 0) {
                ptx::prefetch_tensormap(&maps.us);
                ptx::prefetch_tensormap(&maps.up);
                ptx::prefetch_tensormap(&maps.ut);
                ptx::prefetch_tensormap(&maps.un);
                ptx::prefetch_tensormap(&maps.wc);
            }
        }
        if (warp == 0)
            tm::alloc(tslot, 512);
        tm::fence_before();
        __syncthreads();
        tm::fence_after();
        ptx::pdl_launch_dependents(); // the next kernel of the stream may be scheduled as this one's CTAs retire
        ptx::pdl_wait();              // set-up above ran under the previous kernel's tail; its data is visible from here
        if (p.gate.wait_flag) { // the halo of wcon is being unpacked by a kernel on another stream
            if (threadIdx.x == 0) {
                ptx::gate_wait(p.gate.wait_flag, p.gate.wait_value, p.gate.timeouts);
                ptx::fence_proxy_async_all();
            }
            __syncthreads();
        }
        const uint32_t tbase = *tslot;
        const uint32_t tw = tbase + ((uint32_t)((pw & 3) * 32) << 16) + (uint32_t)((pw >> 2) * 256);

        const int nk = p.nk;
        const T dtr = p.dtr;
        const uint64_t pol_keep = ptx::policy_evict_last(), pol_stream = ptx::policy_evict_first();
        const int total = gridDim.x * PAIRS;
        const int nchunks = (nk + KC - 1) / KC;
        const int c_tail = (nk - 1) / KC;   // first forward chunk that needs per-level checks (holds level nk-1)
        const int c_split = p.k_split / KC; // physical chunks [0, c_split) live in TMEM, the rest in the slab
        const int cb_top = (nk - 2) / KC;   // last stored chunk (holds level nk-2)
        const int nstored = cb_top + 1;
        const int bar_id = 1 + pw;
        // diagnosis only (va.debug & 128): globaltimer stamps per pair -- [0] start, F warp pass p: [1 + 4p] begin,
        // [2 + 4p] end; B warp pass p: [3 + 4p] begin, [4 + 4p] end
        long long *const trace = (p.debug & 128) && lane == 0 && !idle
                                     ? p.trace + (size_t)(blockIdx.x * 8 + pw) * 32
                                     : nullptr;
        auto stamp = [&](int e) {
            if (trace && e < 32)
                trace[e] = (long long)globaltimer_ns();
        };
        if (!is_b)
            stamp(0);

        if (idle) {
            // a pair slot left empty so that the strips of a launch divide evenly among the pairs
        } else if (!is_b) {
            // ============================================================ F warp: forward sweeps back to back
            int f_issue = 0, f_wait = 0;
            uint32_t f_phase = 0;
            auto issue_f = [&](int i0, int j, int c) {
                const int s = f_issue;
                f_issue = f_issue + 1 == S ? 0 : f_issue + 1;
                if (ptx::elect_one()) {
                    unsigned char *st = fring + s * fstage;
                    uint64_t *bar = &ffull[s];
                    ptx::mbar_expect_tx(bar, ftx);
                    ptx::tma_load_3d_hint(st, &maps.us, bar, i0, j, c * KC, pol_stream);
                    ptx::tma_load_3d_hint(st + KC * 32 * es, &maps.up, bar, i0, j, c * KC, pol_keep); // B warp re-reads
                    ptx::tma_load_3d_hint(st + 2 * KC * 32 * es, &maps.ut, bar, i0, j, c * KC, pol_stream);
                    ptx::tma_load_3d_hint(st + 3 * KC * 32 * es, &maps.un, bar, i0, j, c * KC + 1, pol_stream);
                    ptx::tma_load_3d_hint(st + 4 * KC * 32 * es, &maps.wc, bar, i0, j, c * KC + 1, pol_stream);
                }
            };
            auto prime = [&](int item, T &u0) { // first S-1 forward chunks of a strip + u_stage(k = 0)
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i0 = ti * 32;
                for (int c = 0; c < S - 1 && c < nchunks; ++c)
                    issue_f(i0, j, c);
                u0 = i0 + lane < p.ni ? __ldg(p.u_stage.ptr + i0 + lane + (int64_t)j * p.u_stage.sj) : T(0);
            };
            // first strips: the pairs of a CTA take neighbouring strips (va.debug & 256, experiment: the CTA then reads 2 KB
            // contiguous rows per level and field at about the same time) or strips one grid apart
            int item = (p.debug & 256) ? blockIdx.x * PAIRS + pw : pw * gridDim.x + blockIdx.x;
            T u0 = T(0);
            if (item < p.items)
                prime(item, u0);
            int mirrored = 0, pass = 0;
            if (item >= p.items) { // no strip for this pair: release the partner
                if (lane == 0)
                    ctl[2] = -1;
                __syncwarp();
                named_barrier_sync(bar_id, 64);
            }
            while (item < p.items) {
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i0 = ti * 32;
                va_state<T> st;
                st.u_k = u0;
                st.u_km1 = st.wsum_k = st.cc_prev = st.dc_prev = st.up_last = T(0);
                const int need0 = (pass - 1) * nstored; // chunks the partner had read before this pass
                stamp(1 + 4 * pass);
                for (int c = 0; c < nchunks; ++c) {
                    if (c + S - 1 < nchunks) // refill the stage consumed in the previous iteration
                        issue_f(i0, j, c + S - 1);
                    const int s = f_wait;
                    ptx::mbar_wait(&ffull[s], f_phase);
                    if (++f_wait == S) {
                        f_wait = 0;
                        f_phase ^= 1;
                    }
                    const T *sd = reinterpret_cast<const T *>(fring + s * fstage);
                    const T *wc = sd + 4 * KC * 32;
                    T ccv[KC], dcv[KC];
                    if (c != 0 && c < c_tail) {
                        va_forward_body_chunk<T, KC>(
                            dtr, st,
                            [&](int u, T &us, T &un, T &w0, T &w1, T &up, T &ut) {
                                us = sd[u * 32 + lane], up = sd[(KC + u) * 32 + lane], ut = sd[(2 * KC + u) * 32 + lane];
                                un = sd[(3 * KC + u) * 32 + lane];
                                w0 = wc[u * L::ww + lane], w1 = wc[u * L::ww + lane + 1];
                            },
                            [&](int u, T cc, T dc, T) {
                                ccv[u] = cc;
                                dcv[u] = dc;
                            });
                    } else {
#pragma unroll
                        for (int u = 0; u < KC; ++u) {
                            const int k = c * KC + u;
                            ccv[u] = dcv[u] = T(0);
                            if (k < nk) {
                                T us = sd[u * 32 + lane], up = sd[(KC + u) * 32 + lane], ut = sd[(2 * KC + u) * 32 + lane];
                                T un = sd[(3 * KC + u) * 32 + lane];
                                T w0 = wc[u * L::ww + lane], w1 = wc[u * L::ww + lane + 1];
                                va_forward_level<T>(k, nk, dtr, us, un, w0, w1, up, ut, st, ccv[u], dcv[u]);
                            }
                        }
                    }
                    if (c <= cb_top) {
                        const int ph = mirrored ? cb_top - c : c;
                        if (pass > 0) { // the partner must have read this slot (chunk cb_top - c of the previous strip)
                            while (ctl[0] < need0 + c + 1) {
                            }
                            if (ph < c_split)
                                tm::fence_after();
                            else
                                __threadfence_block();
                        }
                        if (ph < c_split) {
                            uint32_t w[NW];
#pragma unroll
                            for (int u = 0; u < KC; ++u)
                                tm::pack<T>(ccv[u], dcv[u], w + u * CPL);
                            tm::st<NW>(tw + (uint32_t)(ph * NW), w);
                        } else {
                            T *q = ss + (ph - c_split) * (KC * 64);
#pragma unroll
                            for (int u = 0; u < KC; ++u) {
                                if (c * KC + u < nk - 1) {
                                    q[u * 64] = ccv[u];
                                    q[u * 64 + 32] = dcv[u];
                                }
                            }
                        }
                    }
                    __syncwarp(); // all lanes are done with stage s before one lane refills it
                }
                // ---- hand the strip to the partner and go on
                stamp(2 + 4 * pass);
                const int par = pass & 1;
                hand[(par * 2) * 32] = st.dc_prev;
                hand[(par * 2 + 1) * 32] = st.up_last;
                int next = 0;
                if (lane == 0)
                    next = atomicAdd(p.tickets, 1);
                next = total + __shfl_sync(0xffffffffu, next, 0);
                if (lane == 0) {
                    ctl[2 + par * 3] = item;
                    ctl[3 + par * 3] = next < p.items;
                    ctl[4 + par * 3] = mirrored;
                }
                if (next < p.items)
                    prime(next, u0);
                tm::wait_st();
                tm::fence_before();
                __syncwarp();
                named_barrier_sync(bar_id, 64);
                item = next;
                mirrored ^= 1;
                ++pass;
            }
        } else {
            // ============================================================ B warp: backward sweep of the partner's last strip
            int b_issue = 0, b_wait = 0;
            uint32_t b_phase = 0;
            const int64_t us_sk = p.utens_stage.sk;
            int done = 0; // chunks read so far (published in ctl[0])
            for (int pass = 0;; ++pass) {
                named_barrier_sync(bar_id, 64);
                tm::fence_after();
                const int par = pass & 1;
                const int item = ctl[2 + par * 3];
                if (item < 0)
                    break;
                const int more = ctl[3 + par * 3], mirrored = ctl[4 + par * 3];
                stamp(3 + 4 * pass);
                T data = hand[(par * 2) * 32];
                const T up_last = hand[(par * 2 + 1) * 32];
                const int ti = item % p.tiles_i, j = item / p.tiles_i;
                const int i0 = ti * 32;
                const bool active = i0 + lane < p.ni;
                if constexpr (BLDG) {
                    va_pair_backward_ldg<T, KC>(p, lane, i0, j, cb_top, c_split, mirrored, more, tw, ss, ctl, data, up_last,
                        done, pol_stream);
                } else {
                    auto issue_b = [&](int cb) { // u_pos of chunk cb: an L2 hit (the forward load left the lines evict_last)
                        const int s = b_issue;
                        b_issue = b_issue + 1 == SB ? 0 : b_issue + 1;
                        if (ptx::elect_one()) {
                            ptx::mbar_expect_tx(&bfull[s], btx);
                            ptx::tma_load_3d_hint(bring + s * bstage, &maps.up, &bfull[s], i0, j, cb * KC, pol_stream);
                        }
                    };
                    for (int n = 0; n < SB - 1 && cb_top - n >= 0; ++n)
                        issue_b(cb_top - n);
                    T *o = p.utens_stage.ptr + (active ? i0 + lane : 0) + (int64_t)j * p.utens_stage.sj + (int64_t)(nk - 1) * us_sk;
                    if (active) // last_level :118-121
                        *o = dtr * (data - up_last);
                    // The k-cache is read one chunk ahead of the arithmetic: the TMEM / shared-memory latency of chunk cb-1
                    // hides behind the four levels of chunk cb.  While the partner sweeps forward (`more`) it is told
                    // after every second chunk which slots are free; in the last, backward-only pass nobody listens.
                    T bcc[KC], bdc[KC];
                    uint32_t w[NW];
                    auto fetch = [&](int cb) { // issue the loads of chunk cb
                        const int ph = mirrored ? cb_top - cb : cb; // where the strip stored its chunk cb
                        if (ph < c_split)
                            tm::ld<NW>(tw + (uint32_t)(ph * NW), w);
                        return ph;
                    };
                    auto land = [&](int cb, int ph) { // ... and wait for them
                        if (ph < c_split) {
                            tm::wait_ld();
    #pragma unroll
                            for (int u = 0; u < KC; ++u)
                                tm::unpack<T>(w + u * CPL, bcc[u], bdc[u]);
                        } else {
                            const T *q = ss + (ph - c_split) * (KC * 64);
    #pragma unroll
                            for (int u = 0; u < KC; ++u) {
                                bcc[u] = bdc[u] = T(0);
                                if (cb * KC + u < nk - 1) {
                                    bcc[u] = q[u * 64];
                                    bdc[u] = q[u * 64 + 32];
                                }
                            }
                        }
                        ++done;
                        if (more && ((done & 1) == 0 || cb == 0)) { // the slots read so far may be overwritten by the partner
                            tm::fence_before();
                            __threadfence_block();
                            __syncwarp();
                            if (lane == 0)
                                ctl[0] = done;
                        }
                    };
                    auto sweep = [&](int cb, auto checked) { // levels of chunk cb, top down (body :111-116)
                        T ccv[KC], dcv[KC];
    #pragma unroll
                        for (int u = 0; u < KC; ++u)
                            ccv[u] = bcc[u], dcv[u] = bdc[u];
                        int ph_next = 0;
                        if (cb > 0)
                            ph_next = fetch(cb - 1);
                        if (cb - (SB - 1) >= 0)
                            issue_b(cb - (SB - 1));
                        const int s = b_wait;
                        ptx::mbar_wait(&bfull[s], b_phase);
                        if (++b_wait == SB) {
                            b_wait = 0;
                            b_phase ^= 1;
                        }
                        const T *sb = reinterpret_cast<const T *>(bring + s * bstage) + lane;
    #pragma unroll
                        for (int u = KC - 1; u >= 0; --u) {
                            if (!decltype(checked)::value || cb * KC + u <= nk - 2) {
                                data = dcv[u] - ccv[u] * data;
                                o -= us_sk;
                                if (active)
                                    *o = dtr * (data - sb[u * 32]);
                            }
                        }
                        if (cb > 0)
                            land(cb - 1, ph_next);
                        __syncwarp(); // all lanes are done with the ring stage before one lane refills it
                    };
                    land(cb_top, fetch(cb_top));
                    sweep(cb_top, std::true_type()); // the top chunk may hold fewer than KC levels
                    for (int cb = cb_top - 1; cb >= 0; --cb)
                        sweep(cb, std::false_type());
                }
                stamp(4 + 4 * pass);
                if (!more)
                    break;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(p.tickets + 1, 1) == (int)gridDim.x - 1) { // every F warp has drawn its last ticket
                p.tickets[0] = 0;
                p.tickets[1] = 0;
                __threadfence();
                if (p.gate.post)
                    atomicAdd(p.gate.post, 1ULL); // the launch is done: a later unpack may overwrite the halos it read
            }
        }
        tm::fence_before();
        __syncthreads();
        if (warp == 0)
            tm::dealloc(tbase, 512);
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

    template <class T, int KC, int S, int NB, int NS>
    int launch_va_stream(const va_maps &maps, const va_params<T> &p, int grid, cudaStream_t stream) {
        auto kernel = va_stream_kernel<T, KC, S, NB, NS>;
        const int smem = va_stream_smem<T, KC, S>();
        static thread_local int done_dev = -1;
        if (done_dev != dev()->device) {
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done_dev = dev()->device;
        }
        kernel<<<grid, 32, smem, stream>>>(maps, p);
        count_launch();
        return check_launch("va_stream_kernel");
    }

    // va.unroll = levels per TMA stage (4 default, 8), va.stages = ring depth, va.threads (re-used) = register blocks
    // of 8 levels the backward sweep keeps in flight (32 -> 1 ... 128 -> 4; default 4)
    template <class T, int NS>
    int dispatch_va_stream(const va_maps &maps, const va_params<T> &p, int kc, int stages, int nb, int grid,
        cudaStream_t stream) {
        if (kc == 8)
            return launch_va_stream<T, 8, 3, 4, NS>(maps, p, grid, stream);
        if (stages == 2) // shallow rings: room for 14 warps per SM (one round of strips)
            return nb == 1 ? launch_va_stream<T, 4, 2, 1, NS>(maps, p, grid, stream)
                           : launch_va_stream<T, 4, 2, 2, NS>(maps, p, grid, stream);
        switch (nb) {
        case 1:
            return launch_va_stream<T, 4, 4, 1, NS>(maps, p, grid, stream);
        case 2:
            return launch_va_stream<T, 4, 4, 2, NS>(maps, p, grid, stream);
        case 3:
            return launch_va_stream<T, 4, 4, 3, NS>(maps, p, grid, stream);
        default:
            break;
        }
        switch (stages) {
        case 3:
            return launch_va_stream<T, 4, 3, 4, NS>(maps, p, grid, stream);
        case 6:
            return launch_va_stream<T, 4, 6, 4, NS>(maps, p, grid, stream);
        default:
            return launch_va_stream<T, 4, 4, 4, NS>(maps, p, grid, stream);
        }
    }

    constexpr int kTraceEvents = 32, kTraceSlots = 2048;
    constexpr size_t kTraceBytes = (size_t)kTraceSlots * kTraceEvents * sizeof(long long);

    // {next strip, finished warps} of the TMEM variants, zeroed once (every launch resets its own pair when it ends).
    // Launches rotate over 16 pairs, so launches that run at the same time on different streams do not share one.
    int *va_ticket_base();
    std::atomic<unsigned> g_va_launch{0};
    int *va_ticket_counters() {
        int *base = va_ticket_base();
        return base ? base + 2 * (g_va_launch.fetch_add(1, std::memory_order_relaxed) % 16) : nullptr;
    }
    int *va_ticket_base() {
        static int *ctr[64] = {};
        static std::mutex mtx; // first use may come from several host threads
        std::lock_guard<std::mutex> lock(mtx);
        const int d = dev()->device;
        if (d < 0 || d >= 64)
            return nullptr;
        if (!ctr[d]) {
            int *q = nullptr; // 256 bytes of counters + the (diagnosis only) time-stamp trace of va.debug & 128
            if (cudaMalloc(&q, 256 + kTraceBytes) != cudaSuccess || cudaMemset(q, 0, 256 + kTraceBytes) != cudaSuccess) {
                cuda_fail(cudaGetLastError(), "va ticket counters");
                return nullptr;
            }
            ctr[d] = q;
        }
        return ctr[d];
    }

    // TMEM variant: one CTA per SM.  va.ctas_per_sm = warps per CTA (4..8, default 8; < 0: an absolute number of
    // CTAs of 8 warps, tests), va.stages = TMA ring depth (0: as deep as shared memory allows, at most 8),
    // va.threads (re-used) = start stagger between the warps of a CTA in units of 100 ns.  Returns -1 when the
    // shared-memory slab for the levels above the TMEM capacity does not fit (tall columns): the caller then uses
    // variant 3.
    template <class T>
    int vert_adv_pair(va_params<T> &p, const options &o, device_state *d, const va_maps &maps, cudaStream_t stream) {
        constexpr int KC = 4;
        int pairs = o.va_ctas_per_sm >= 1 && o.va_ctas_per_sm <= 8 ? o.va_ctas_per_sm : 0;
        const int64_t strips = (int64_t)p.tiles_i * p.nj;
        if (pairs == 0)
            pairs = 8;
        p.pairs = pairs;
        p.debug = o.va_debug;
        const int levels = va_tmem_cfg<T>::levels(8);
        const int cb_top = (p.nk - 2) / KC;
        int k_split = (cb_top + 1) * KC;
        if (k_split > levels)
            k_split = levels;
        p.k_split = k_split;
        const int slab_bytes = (cb_top + 1 - k_split / KC) * KC * 64 * (int)sizeof(T);
        p.slots = slab_bytes;
        p.scratch = nullptr;
        p.persistent = 1;
        const int bstages = o.va_unroll >= 2 && o.va_unroll <= 8 ? o.va_unroll : 3;
        int stages = o.va_stages >= 2 && o.va_stages <= 16 ? o.va_stages : 0;
        if (stages == 0)
            for (stages = 6; stages > 2 && va_pair_smem<T, KC>(pairs, stages, bstages, slab_bytes) > d->max_smem_optin;)
                --stages;
        const int smem = va_pair_smem<T, KC>(pairs, stages, bstages, slab_bytes);
        if (smem > d->max_smem_optin)
            return -1;
        p.stages = stages;
        p.bstages = bstages;
        p.stagger_ns = 0;
        p.tickets = va_ticket_counters();
        p.trace = reinterpret_cast<long long *>(reinterpret_cast<char *>(va_ticket_base()) + 256);
        if (!p.tickets)
            return GTB_ERR_ALLOC;
        int grid = o.va_ctas_per_sm < 0 ? -o.va_ctas_per_sm : stencil_sms(d);
        if ((int64_t)grid * pairs > strips)
            grid = (int)((strips + pairs - 1) / pairs);
        int st = set_l2_persist(0);
        if (st)
            return st;
        const bool bldg = o.va_bldg != 2; // 0 auto / 1: u_pos of the backward sweep by register loads; 2: through a TMA ring
        auto kernel = bldg ? va_pair_kernel<T, KC, true> : va_pair_kernel<T, KC, false>;
        static thread_local int done_dev[2] = {-1, -1};
        if (done_dev[bldg] != d->device) {
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, d->max_smem_optin));
            GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            done_dev[bldg] = d->device;
        }
        GTB_CUDA(launch_pdl(kernel, dim3(grid), dim3(512), (size_t)smem, stream, maps, p));
        count_launch();
        return check_launch("va_pair_kernel");
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

    // Persistent one-warp CTAs over the 32-column strips.  va.variant 2: per-warp TMA ring + per-thread slab loads;
    // 3 (and auto): two-ring streaming kernel.  va.ctas_per_sm > 0: warps per SM, < 0: absolute grid size (tests).
    // The TMA kernels: va.variant 0 (auto) = the paired-warp TMEM kernel for fp64 (fastest measured, profiles/README.md)
    // and, for fp32 -- where a strip moves half the bytes and the sweep is latency-bound -- and for columns too tall
    // for TMEM + shared memory, the one-warp streaming kernel with its L2 slab (variant 3).  Variants 2, 4, 5 and 6 of
    // round 1 (first TMA version, register-resident k cache, one-warp TMEM, fused-sweep TMEM) were the measured-slower
    // steps towards variant 7 and have been removed.
    template <class T>
    int vert_adv_tma(va_params<T> &p, const options &o, device_state *d, cudaStream_t stream, bool *done) {
        *done = false;
        if (o.va_variant == 2 || (o.va_variant >= 4 && o.va_variant <= 6))
            return fail(GTB_ERR_ARG, "gtb_vert_adv: va.variant=%d was removed (use 0 = auto, 1 = LDG fallback, 3 = "
                                     "streaming warps with an L2 slab, 7 = paired warps with the k cache in TMEM)", o.va_variant);
        const int tm_variant = o.va_variant == 0 ? (sizeof(T) == 8 ? 7 : 3) : o.va_variant;
        const int kc = 4; // levels per TMA stage
        p.kc = kc;
        p.debug = o.va_debug;
        va_maps maps;
        if (!make_va_maps<T>(maps, p))
            return GTB_OK; // not addressable: the caller falls back to the register-prefetch kernel
        const int wps = o.va_ctas_per_sm > 0 ? o.va_ctas_per_sm : 7; // warps per SM (variant 3)
        const int64_t strips = (int64_t)p.tiles_i * p.nj;
        p.items = (int)strips;
        int grid = o.va_ctas_per_sm < 0 ? -o.va_ctas_per_sm : wps * stencil_sms(d);
        if (grid > strips)
            grid = (int)strips;
        if (tm_variant != 7 && (p.gate.wait_flag || p.gate.post))
            return fail(GTB_ERR_ARG, "gtb_vert_adv: a gate needs the paired-warp kernel");
        if (tm_variant == 7) {
            int st = vert_adv_pair<T>(p, o, d, maps, stream);
            if (st >= 0) {
                *done = true;
                return st;
            }
            if (p.gate.wait_flag || p.gate.post)
                return fail(GTB_ERR_ARG, "gtb_vert_adv: a gate needs the paired-warp kernel, which cannot hold nk = %d levels", p.nk);
        }
        const bool save_upos = o.va_save_upos != 2;
        const int ns = save_upos ? 3 : 2;
        p.slots = (int64_t)p.nk * ns * 32; // one contiguous slab per warp
        const int64_t slab_elems = p.slots * grid;
        p.scratch = static_cast<T *>(scratch((size_t)slab_elems * sizeof(T), stream));
        if (!p.scratch)
            return GTB_ERR_ALLOC;
        p.persistent = 1;
        *done = true;
        {
            const int64_t slab = slab_elems * (int64_t)sizeof(T);
            int st = set_l2_persist(o.l2_persist_mb < 0 ? slab : (int64_t)o.l2_persist_mb << 20);
            if (st)
                return st;
        }
        return save_upos ? dispatch_va_stream<T, 3>(maps, p, kc, o.va_stages, o.va_threads / 32, grid, stream)
                         : dispatch_va_stream<T, 2>(maps, p, kc, o.va_stages, o.va_threads / 32, grid, stream);
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
        p.gate = take_gate();
        if ((p.gate.wait_flag || p.gate.post) && !(sizeof(T) == 8 && (o.va_variant == 0 || o.va_variant == 7)))
            return fail(GTB_ERR_ARG, "gtb_vert_adv: a gate (gtb_stencil_gate) needs the paired-warp fp64 kernel (va.variant 0 or 7)");
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
        p.scratch = static_cast<T *>(scratch((size_t)ns * nk * p.slots * sizeof(T), s));
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

GTB_API int gtb_debug_trace(void *dst, int64_t bytes) {
    if (!dst || bytes < 0)
        return fail(GTB_ERR_ARG, "gtb_debug_trace: bad argument");
    if (!dev())
        return GTB_ERR_CUDA;
    int *base = va_ticket_base();
    if (!base)
        return GTB_ERR_ALLOC;
    if ((size_t)bytes > kTraceBytes)
        bytes = (int64_t)kTraceBytes;
    GTB_CUDA(cudaDeviceSynchronize());
    GTB_CUDA(cudaMemcpy(dst, reinterpret_cast<char *>(base) + 256, (size_t)bytes, cudaMemcpyDeviceToHost));
    GTB_CUDA(cudaMemset(reinterpret_cast<char *>(base) + 256, 0, kTraceBytes));
    return GTB_OK;
}
