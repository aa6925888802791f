// hori_diff.cu -- fused horizontal diffusion (horizontal_diffusion.cpp:35-106) for sm_100a.
//
// What the reference does (stencil/gpu/launch_kernel.hpp:83-119, make_kernel_fun.hpp:101-125): one CTA of 64x12
// threads per 64x8 tile and per k level, lap/flx/fly go through three shared-memory ij_caches separated by
// __syncthreads, `in` is read ~10x per point through L1 and never staged.
//
// What this kernel does instead:
//  * persistent CTAs (grid = SMs x resident CTAs) walk a flat list of (tile, k) work items, so the 256x256x80 case
//    is 5120 items spread evenly over 148 SMs instead of 10240 short-lived CTAs;
//  * the halo-extended `in` tile (BI+4)x(BJ+4) and the coeff tile BIxBJ of each item are brought to shared memory by
//    TMA (cp.async.bulk.tensor.3d, one elected thread, completion on an mbarrier) through a STAGES-deep ring, so
//    every SM keeps several tiles of HBM traffic in flight without spending registers or LSU issue slots on it;
//    out-of-domain parts of edge tiles are zero-filled by the TMA unit (no out-of-bounds reads);
//  * lap, flx and fly never touch shared memory: each thread owns a column of R outputs and keeps the 5-wide
//    column neighbourhood in registers (the ij_caches of the reference become register tiles), which removes all
//    three block-wide barriers of the reference and leaves exactly one __syncthreads per item (ring hand-over);
//  * `out` is stored straight from registers, 256 B (fp64) per warp instruction.
// A second variant stages the same tiles with cp.async (LDGSTS) element by element; it is used when the layout
// does not satisfy TMA's 16-byte stride/alignment rules, and serves as an A/B baseline ("hd.variant" option: 0 auto, 1 cp.async, 2 TMA + block barrier).
//
// Arithmetic follows the functor bodies operation by operation (no FMA contraction: the file is compiled with
// -fmad=false), so results are bit-identical to oracle/gt_oracle.c compiled with -ffp-contract=off.
#include "common.cuh"
#include "tma.cuh"

using namespace gtb;

namespace {

    constexpr int BI = 64;  // tile extent along i (one warp row = 32 consecutive i)
    constexpr int BJ = 16;  // tile extent along j
    constexpr int R = 4;    // outputs per thread (consecutive j)
    constexpr int THREADS = BI * (BJ / R); // compute threads
    constexpr int WARPS = THREADS / 32;
    constexpr int IN_H = BJ + 4;

    template <class T>
    struct layout {
        // The staged `in` tile starts LEAD elements before the tile so that every TMA box starts on a 16-byte
        // boundary (the innermost box coordinate must be 16-byte aligned): 2 doubles or 4 floats; only 2 are used.
        static constexpr int lead = 16 / (int)sizeof(T);
        static constexpr int in_w = BI + 2 * lead;
        static constexpr int in_bytes = in_w * IN_H * (int)sizeof(T);
        static constexpr int co_bytes = BI * BJ * (int)sizeof(T);
        static constexpr int in_alloc = (in_bytes + 127) / 128 * 128;
        static constexpr int co_alloc = (co_bytes + 127) / 128 * 128;
        static constexpr int stage_bytes = in_alloc + co_alloc;
    };

    template <class T>
    struct hd_params {
        const T *in;
        const T *coeff;
        T *out;
        int64_t in_sj, in_sk, co_sj, co_sk, out_sj, out_sk;
        int ni, nj, nk;
        int tiles_i, tiles_j;
        int step_i, step_j, step_k; // gridDim.x decomposed in (tile_i, tile_j, k) digits
        stencil_gate gate;          // device-side ordering against a concurrent halo exchange (empty: none)
        const T *crlato, *crlatu;   // simple_hori_diff only: j-only coefficients (compute-domain j = 0), element strides
        int64_t cro_sj, cru_sj;
    };

    // Position of a CTA in the flat item list (i fastest, then j, then k: CTAs that run side by side work on
    // neighbouring tiles of the same level, so the halo rows they share are still in L2).  Advancing by gridDim.x
    // items is three adds with carry instead of 64-bit divisions.
    struct item_iter {
        int ti, tj, k;
        template <class P>
        __device__ __forceinline__ void start(const P &p, int w) {
            ti = w % p.tiles_i;
            int r = w / p.tiles_i;
            tj = r % p.tiles_j;
            k = r / p.tiles_j;
        }
        template <class P>
        __device__ __forceinline__ void next(const P &p) {
            ti += p.step_i;
            int c = ti >= p.tiles_i;
            ti -= c * p.tiles_i;
            tj += p.step_j + c;
            c = tj >= p.tiles_j;
            tj -= c * p.tiles_j;
            k += p.step_k + c;
        }
        __device__ __forceinline__ int i0() const { return ti * BI; }
        __device__ __forceinline__ int j0() const { return tj * BJ; }
    };

    // The four stages for one thread: column i = tx, rows j0+ty*R .. +R-1, reading the staged tiles.  `release` is
    // called once the thread's shared-memory reads are done (the stage can be refilled while the math runs).
    template <class T, bool SIMPLE = false, class Release>
    __device__ __forceinline__ void compute_item(const hd_params<T> &p, const T *__restrict__ sin,
        const T *__restrict__ sco, const item_iter &it, int tx, int ty, Release &&release) {
        constexpr int W = layout<T>::in_w;
        const int jl = ty * R;
        const T *c = sin + (jl + 2) * W + tx + layout<T>::lead; // in(i, j0+jl)
        T c0[R + 4], cp1[R + 2], cm1[R + 2], cp2[R], cm2[R];
#pragma unroll
        for (int d = 0; d < R + 4; ++d)
            c0[d] = c[(d - 2) * W];
#pragma unroll
        for (int d = 0; d < R + 2; ++d) {
            cp1[d] = c[(d - 1) * W + 1];
            cm1[d] = c[(d - 1) * W - 1];
        }
#pragma unroll
        for (int d = 0; d < R; ++d) {
            cp2[d] = c[d * W + 2];
            cm2[d] = c[d * W - 2];
        }
        T co[R];
#pragma unroll
        for (int d = 0; d < R; ++d)
            co[d] = sco[(jl + d) * BI + tx];
        release();

        if constexpr (SIMPLE) {
            // simple_hori_diff.cpp:25-61 on the same staged neighbourhood.  wlap_function: in(1,0) + in(-1,0) - 2 in +
            // crlato (in(0,1) - in) + crlatu (in(0,-1) - in); divflux_function: in + ((fluxx_m - fluxx) +
            // (fluxy_m - fluxy)) coeff.  crlato / crlatu depend on j only.
            T cro[R + 2], cru[R + 2]; // rows j = -1 .. R
#pragma unroll
            for (int d = 0; d < R + 2; ++d) {
                int gj = it.j0() + jl - 1 + d;
                gj = gj > p.nj + 1 ? p.nj + 1 : gj; // rows past the domain only feed masked outputs
                cro[d] = p.crlato[(int64_t)gj * p.cro_sj];
                cru[d] = p.crlatu[(int64_t)gj * p.cru_sj];
            }
            T lap_c[R + 2], lap_p[R], lap_m[R];
#pragma unroll
            for (int d = 0; d < R + 2; ++d) {
                const T cc = c0[d + 1];
                lap_c[d] = cp1[d] + cm1[d] - T(2) * cc + cro[d] * (c0[d + 2] - cc) + cru[d] * (c0[d] - cc);
            }
#pragma unroll
            for (int d = 0; d < R; ++d) {
                const T cpp = cp1[d + 1], cmm = cm1[d + 1];
                lap_p[d] = cp2[d] + c0[d + 2] - T(2) * cpp + cro[d + 1] * (cp1[d + 2] - cpp) + cru[d + 1] * (cp1[d] - cpp);
                lap_m[d] = c0[d + 2] + cm2[d] - T(2) * cmm + cro[d + 1] * (cm1[d + 2] - cmm) + cru[d + 1] * (cm1[d] - cmm);
            }
            const int i = it.i0() + tx;
            const int j = it.j0() + jl;
            T *o = p.out + i + (int64_t)j * p.out_sj + (int64_t)it.k * p.out_sk;
#pragma unroll
            for (int d = 0; d < R; ++d) {
                const T lc = lap_c[d + 1];
                const T fluxx = lap_p[d] - lc;
                const T fluxx_m = lc - lap_m[d];
                const T fluxy = cro[d + 1] * (lap_c[d + 2] - lc);
                const T fluxy_m = cro[d + 1] * (lc - lap_c[d]);
                const T res = c0[d + 2] + ((fluxx_m - fluxx) + (fluxy_m - fluxy)) * co[d];
                if (i < p.ni && j + d < p.nj)
                    o[(int64_t)d * p.out_sj] = res;
            }
            return;
        }

        // lap_function (horizontal_diffusion.cpp:35-47): 4*in - (in(1,0) + in(0,1) + in(-1,0) + in(0,-1))
        T lap_c[R + 2]; // lap(i, j) for j = -1 .. R
#pragma unroll
        for (int d = 0; d < R + 2; ++d)
            lap_c[d] = T(4) * c0[d + 1] - (cp1[d] + c0[d + 2] + cm1[d] + c0[d]);
        T lap_p[R], lap_m[R]; // lap(i+1, j), lap(i-1, j) for j = 0 .. R-1
#pragma unroll
        for (int d = 0; d < R; ++d) {
            lap_p[d] = T(4) * cp1[d + 1] - (cp2[d] + cp1[d + 2] + c0[d + 2] + cp1[d]);
            lap_m[d] = T(4) * cm1[d + 1] - (c0[d + 2] + cm1[d + 2] + cm2[d] + cm1[d]);
        }
        // fly_function (:63-75) at j = -1 .. R-1
        T fly[R + 1];
#pragma unroll
        for (int d = 0; d < R + 1; ++d) {
            T res = lap_c[d + 1] - lap_c[d];
            fly[d] = res * (c0[d + 2] - c0[d + 1]) > T(0) ? T(0) : res;
        }
        const int i = it.i0() + tx;
        const int j = it.j0() + jl;
        T *o = p.out + i + (int64_t)j * p.out_sj + (int64_t)it.k * p.out_sk;
#pragma unroll
        for (int d = 0; d < R; ++d) {
            // flx_function (:49-61) at (i, j) and (i-1, j)
            T rx = lap_p[d] - lap_c[d + 1];
            T flx = rx * (cp1[d + 1] - c0[d + 2]) > T(0) ? T(0) : rx;
            T rxm = lap_c[d + 1] - lap_m[d];
            T flxm = rxm * (c0[d + 2] - cm1[d + 1]) > T(0) ? T(0) : rxm;
            // out_function (:77-91)
            T res = c0[d + 2] - co[d] * (flx - flxm + fly[d + 1] - fly[d]);
            if (i < p.ni && j + d < p.nj)
                o[(int64_t)d * p.out_sj] = res;
        }
    }

    template <class T>
    __device__ __forceinline__ void tma_issue(const CUtensorMap *map_in, const CUtensorMap *map_co, unsigned char *base,
        uint64_t *bar, const item_iter &it) {
        using L = layout<T>;
        ptx::mbar_expect_tx(bar, L::in_bytes + L::co_bytes);
        // tensor coordinate 0 of map_in is element (-lead, -2); of map_co element (0, 0)
        ptx::tma_load_3d(base, map_in, bar, it.i0(), it.j0(), it.k);
        ptx::tma_load_3d(base + L::in_alloc, map_co, bar, it.i0(), it.j0(), it.k);
    }

    // ------------------------------------------------ TMA variant with a block barrier per item (hd.variant = 2)
    template <class T, int STAGES, bool SIMPLE = false>
    __global__ void __launch_bounds__(THREADS, 2) hd_tma_kernel(const __grid_constant__ CUtensorMap map_in,
        const __grid_constant__ CUtensorMap map_co, const hd_params<T> p) {
        using L = layout<T>;
        extern __shared__ __align__(128) unsigned char smem[];
        uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * L::stage_bytes);
        const int tid = threadIdx.x;
        const int tx = tid % BI, ty = tid / BI;
        if (tid == 0) {
            ptx::prefetch_tensormap(&map_in);
            ptx::prefetch_tensormap(&map_co);
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                ptx::mbar_init(&full[s], 1);
            ptx::fence_barrier_init();
        }
        // launched with programmatic stream serialization (launch_pdl): everything above ran under the previous kernel's
        // tail; nothing an earlier kernel of the stream may have touched is read or written before this wait
        ptx::pdl_launch_dependents();
        ptx::pdl_wait();
        if (tid == 0 && p.gate.wait_flag) { // the halo of `in` is being unpacked by a kernel on another stream
            ptx::gate_wait(p.gate.wait_flag, p.gate.wait_value, p.gate.timeouts);
            ptx::fence_proxy_async_all(); // ... and is read through the async proxy (TMA) below
        }
        __syncthreads();
        item_iter it, ahead;
        it.start(p, (int)blockIdx.x);
        ahead = it;
        if (tid == 0) {
            for (int s = 0; s < STAGES && ahead.k < p.nk; ++s, ahead.next(p))
                tma_issue<T>(&map_in, &map_co, smem + s * L::stage_bytes, &full[s], ahead);
        }
        for (int n = 0; it.k < p.nk; ++n, it.next(p)) {
            const int s = n % STAGES;
            ptx::mbar_wait(&full[s], (uint32_t)((n / STAGES) & 1));
            const unsigned char *base = smem + s * L::stage_bytes;
            compute_item<T, SIMPLE>(p, reinterpret_cast<const T *>(base), reinterpret_cast<const T *>(base + L::in_alloc),
                it, tx, ty, [] {});
            __syncthreads(); // every thread has consumed stage s: hand it back to the TMA unit
            if (tid == 0 && ahead.k < p.nk) {
                tma_issue<T>(&map_in, &map_co, smem + s * L::stage_bytes, &full[s], ahead);
                ahead.next(p);
            }
        }
        if (p.gate.post) { // tell whoever waits for this launch (the unpack of a later exchange) that it is done
            __threadfence();
            __syncthreads();
            if (tid == 0 && atomicAdd(p.gate.cta_done, 1) == (int)gridDim.x - 1) {
                *p.gate.cta_done = 0;
                __threadfence();
                atomicAdd(p.gate.post, 1ULL);
            }
        }
    }

    // ------------------------------------------------ two tile pipelines in ONE CTA per SM (hd.variant = 4)
    // The same pipeline as hd_tma_kernel, twice per CTA: threads [0, THREADS) and [THREADS, 2*THREADS) each own a TMA
    // ring, mbarriers and a share of the item list, and meet on their own named barrier.  One CTA per SM is what
    // programmatic dependent launch needs here (with two CTAs per SM the early CTAs of the next launch slow the
    // running one down, profiles/r02_hd_pdl.txt): the prologue runs under the previous launch's tail.
    template <class T, int STAGES>
    __global__ void __launch_bounds__(2 * THREADS, 1) hd_tma2_kernel(const __grid_constant__ CUtensorMap map_in,
        const __grid_constant__ CUtensorMap map_co, const hd_params<T> p) {
        using L = layout<T>;
        extern __shared__ __align__(128) unsigned char smem_all[];
        const int sub = threadIdx.x / THREADS, tid = threadIdx.x % THREADS;
        unsigned char *smem = smem_all + sub * (STAGES * L::stage_bytes);
        uint64_t *full = reinterpret_cast<uint64_t *>(smem_all + 2 * STAGES * L::stage_bytes) + sub * STAGES;
        const int tx = tid % BI, ty = tid / BI;
        if (tid == 0) {
            if (sub == 0) {
                ptx::prefetch_tensormap(&map_in);
                ptx::prefetch_tensormap(&map_co);
            }
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                ptx::mbar_init(&full[s], 1);
            ptx::fence_barrier_init();
        }
        ptx::pdl_launch_dependents();
        ptx::pdl_wait(); // nothing an earlier kernel of the stream may have touched is read or written before this
        if (threadIdx.x == 0 && p.gate.wait_flag) { // the halo of `in` is being unpacked by a kernel on another stream
            ptx::gate_wait(p.gate.wait_flag, p.gate.wait_value, p.gate.timeouts);
            ptx::fence_proxy_async_all();
        }
        __syncthreads();
        item_iter it, ahead;
        it.start(p, (int)blockIdx.x * 2 + sub);
        ahead = it;
        if (tid == 0) {
            for (int s = 0; s < STAGES && ahead.k < p.nk; ++s, ahead.next(p))
                tma_issue<T>(&map_in, &map_co, smem + s * L::stage_bytes, &full[s], ahead);
        }
        for (int n = 0; it.k < p.nk; ++n, it.next(p)) {
            const int s = n % STAGES;
            ptx::mbar_wait(&full[s], (uint32_t)((n / STAGES) & 1));
            const unsigned char *base = smem + s * L::stage_bytes;
            compute_item<T, false>(p, reinterpret_cast<const T *>(base), reinterpret_cast<const T *>(base + L::in_alloc), it,
                tx, ty, [] {});
            asm volatile("bar.sync %0, %1;" ::"r"(1 + sub), "n"(THREADS) : "memory"); // this half is done with stage s
            if (tid == 0 && ahead.k < p.nk) {
                tma_issue<T>(&map_in, &map_co, smem + s * L::stage_bytes, &full[s], ahead);
                ahead.next(p);
            }
        }
        if (p.gate.post) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0 && atomicAdd(p.gate.cta_done, 1) == (int)gridDim.x - 1) {
                *p.gate.cta_done = 0;
                __threadfence();
                atomicAdd(p.gate.post, 1ULL);
            }
        }
    }

    // ------------------------------------------------ cp.async (LDGSTS) variant (hd.variant = 1, any alignment)
    template <class T, int STAGES>
    __global__ void __launch_bounds__(THREADS, 2) hd_cpasync_kernel(const hd_params<T> p) {
        using L = layout<T>;
        extern __shared__ __align__(128) unsigned char smem[];
        const int tid = threadIdx.x;
        const int tx = tid % BI, ty = tid / BI;

        auto issue = [&](int s, const item_iter &w) {
            T *sin = reinterpret_cast<T *>(smem + s * L::stage_bytes);
            T *sco = reinterpret_cast<T *>(smem + s * L::stage_bytes + L::in_alloc);
            const T *gin = p.in + (int64_t)w.k * p.in_sk;
            for (int e = tid; e < L::in_w * IN_H; e += THREADS) {
                int r = e / L::in_w, c = e - r * L::in_w;
                int i = w.i0() + c - L::lead, j = w.j0() + r - 2;
                bool ok = i >= -2 && i < p.ni + 2 && j < p.nj + 2; // j >= -2 always
                const T *src = ok ? gin + i + (int64_t)j * p.in_sj : p.in;
                ptx::cp_async<sizeof(T)>(sin + e, src, ok);
            }
            const T *gco = p.coeff + (int64_t)w.k * p.co_sk;
            for (int e = tid; e < BI * BJ; e += THREADS) {
                int r = e / BI, c = e - r * BI;
                int i = w.i0() + c, j = w.j0() + r;
                bool ok = i < p.ni && j < p.nj;
                const T *src = ok ? gco + i + (int64_t)j * p.co_sj : p.coeff;
                ptx::cp_async<sizeof(T)>(sco + e, src, ok);
            }
        };

        item_iter it, ahead;
        it.start(p, (int)blockIdx.x);
        ahead = it;
        for (int s = 0; s < STAGES - 1; ++s) {
            if (ahead.k < p.nk) {
                issue(s, ahead);
                ahead.next(p);
            }
            ptx::cp_async_commit();
        }
        for (int n = 0; it.k < p.nk; ++n, it.next(p)) {
            const int s = n % STAGES;
            ptx::cp_async_wait<STAGES - 2>(); // this thread's copies for item n have landed
            __syncthreads();                  // ... everybody's have, and stage (n-1)%STAGES is free again
            if (ahead.k < p.nk) {
                issue((n + STAGES - 1) % STAGES, ahead);
                ahead.next(p);
            }
            ptx::cp_async_commit();
            const unsigned char *base = smem + s * L::stage_bytes;
            compute_item<T>(p, reinterpret_cast<const T *>(base), reinterpret_cast<const T *>(base + L::in_alloc), it,
                tx, ty, [] {});
        }
        ptx::cp_async_wait<0>();
    }

    // ------------------------------------------------------------------------------------------ host side
    template <class K>
    int prepare_kernel(K kernel, int smem) {
        // one-time per kernel instantiation (the reference does this on every launch, common/cuda_util.hpp:84-88)
        static thread_local K done_for = nullptr;
        static thread_local int done_dev = -1;
        int d = dev()->device;
        if (done_for == kernel && done_dev == d)
            return GTB_OK;
        GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done_for = kernel;
        done_dev = d;
        return GTB_OK;
    }

    template <class T, int STAGES>
    int launch(hd_params<T> &p, int variant, int ctas_per_sm, cudaStream_t stream) {
        using L = layout<T>;
        device_state *d = dev();
        const int smem = STAGES * L::stage_bytes + 16 * STAGES;
        const int64_t items = (int64_t)p.tiles_i * p.tiles_j * p.nk;
        if (items >= (int64_t)1 << 31)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: domain too large (%lld tile-levels)", (long long)items);
        int grid = stencil_sms(d) * ctas_per_sm;
        if (grid > items)
            grid = (int)items;
        p.step_i = grid % p.tiles_i;
        p.step_j = (grid / p.tiles_i) % p.tiles_j;
        p.step_k = (grid / p.tiles_i) / p.tiles_j;
        p.gate = take_gate();
        const bool gated = p.gate.wait_flag || p.gate.post;
        if (gated && p.gate.post && !p.gate.cta_done)
            return GTB_ERR_ALLOC;
        if (variant == 3)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: hd.variant=3 (warp-specialised producer) was removed: it measured "
                                     "slower than the block-barrier pipeline (27.9 against 24.6 us); use 0, 1 or 2");
        if (variant != 1) {
            CUtensorMap map_in, map_co;
            bool ok = make_map<T>(&map_in, p.in, p.in_sj, p.in_sk, L::lead, 2, (int64_t)L::lead + p.ni + 2, p.nj + 4,
                          p.nk, L::in_w, IN_H) &&
                      make_map<T>(&map_co, p.coeff, p.co_sj, p.co_sk, 0, 0, p.ni, p.nj, p.nk, BI, BJ);
            // auto: short launches (<= 32 items per pipeline: 256^2 x 80 has 17) gain the launch gap (24.5 -> 23.3 us), long
            // ones lose (512^2 x 80: 81 -> 89 us; the two-CTA kernel is at 95 % of the HBM peak there): profiles/r02_hd_pdl.txt
            const bool two_in_one = variant == 4 || (variant == 0 && pdl_allowed() && opts().hd_ctas_per_sm == 0 &&
                                                        items <= (int64_t)stencil_sms(d) * 2 * 32);
            if (ok && two_in_one) { // two pipelines per CTA, one CTA per SM, programmatic dependent launch
                int grid2 = stencil_sms(d);
                if ((int64_t)grid2 * 2 > items)
                    grid2 = (int)((items + 1) / 2);
                const int lg = grid2 * 2;
                p.step_i = lg % p.tiles_i;
                p.step_j = (lg / p.tiles_i) % p.tiles_j;
                p.step_k = (lg / p.tiles_i) / p.tiles_j;
                const int smem2 = 2 * smem;
                auto kernel = hd_tma2_kernel<T, STAGES>;
                int st = prepare_kernel(kernel, smem2);
                if (st)
                    return st;
                GTB_CUDA(launch_pdl(kernel, dim3(grid2), dim3(2 * THREADS), (size_t)smem2, stream, map_in, map_co, p));
                count_launch();
                return check_launch("hd_tma2_kernel");
            }
            if (ok) {
                auto kernel = hd_tma_kernel<T, STAGES>;
                int st = prepare_kernel(kernel, smem);
                if (st)
                    return st;
                // Programmatic dependent launch only with one CTA per SM: with two, the early CTAs of the next launch take
                // the slots of the CTAs that finish first and the launch gets slower (stages 3: 24.6 -> 28.6 us); with
                // one CTA per SM and 4-5 stages it is 24.2 us (profiles/r02_hd_pdl.txt).
                GTB_CUDA(launch_pdl_if(pdl_allowed() && ctas_per_sm == 1, kernel, dim3(grid), dim3(THREADS), (size_t)smem, stream,
                    map_in, map_co, p));
                count_launch();
                return check_launch("hd_tma_kernel");
            }
            if (gated)
                return fail(GTB_ERR_ARG, "gtb_hori_diff: a gate (gtb_stencil_gate) needs the default TMA kernel "
                                         "(hd.variant 0 or 2, TMA-addressable layout)");
            if (variant != 0)
                return fail(GTB_ERR_LAYOUT,
                    "gtb_hori_diff: hd.variant=%d (TMA) needs 16-byte aligned origins and stride_j/stride_k that "
                    "are multiples of 16 bytes", variant);
        }
        if (gated)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: a gate (gtb_stencil_gate) needs the default TMA kernel");
        auto kernel = hd_cpasync_kernel<T, STAGES>;
        int st = prepare_kernel(kernel, smem);
        if (st)
            return st;
        kernel<<<grid, THREADS, smem, stream>>>(p);
        count_launch();
        return check_launch("hd_cpasync_kernel");
    }

    template <class T>
    int hori_diff(const gtb_field *in, const gtb_field *coeff, const gtb_field *out, int ni, int nj, int nk,
        void *stream) {
        if (!field_ok(in) || !field_ok(coeff) || !field_ok(out))
            return fail(GTB_ERR_ARG, "gtb_hori_diff: null field");
        if (ni < 0 || nj < 0 || nk < 0)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: negative size");
        if (in->stride_i != 1 || coeff->stride_i != 1 || out->stride_i != 1)
            return fail(GTB_ERR_LAYOUT, "gtb_hori_diff: stride_i must be 1 (i is the unit-stride axis of storage::gpu)");
        if (out->ptr == in->ptr || out->ptr == coeff->ptr)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: out must not alias an input");
        if (!dev())
            return GTB_ERR_CUDA;
        if (ni == 0 || nj == 0 || nk == 0)
            return GTB_OK;
        hd_params<T> p;
        p.in = static_cast<const T *>(in->ptr);
        p.coeff = static_cast<const T *>(coeff->ptr);
        p.out = static_cast<T *>(out->ptr);
        p.in_sj = in->stride_j, p.in_sk = in->stride_k;
        p.co_sj = coeff->stride_j, p.co_sk = coeff->stride_k;
        p.out_sj = out->stride_j, p.out_sk = out->stride_k;
        p.ni = ni, p.nj = nj, p.nk = nk;
        p.tiles_i = ceil_div(ni, BI), p.tiles_j = ceil_div(nj, BJ);
        const options &o = opts();
        int stages = o.hd_stages ? o.hd_stages : 3;
        int ctas = o.hd_ctas_per_sm ? o.hd_ctas_per_sm : 2;
        if (ctas < 1 || ctas > 2)
            return fail(GTB_ERR_ARG, "gtb_hori_diff: hd.ctas_per_sm must be 1 or 2");
        cudaStream_t s = as_stream(stream);
        switch (stages) {
        case 2:
            return launch<T, 2>(p, o.hd_variant, ctas, s);
        case 3:
            return launch<T, 3>(p, o.hd_variant, ctas, s);
        case 4:
            return launch<T, 4>(p, o.hd_variant, ctas, s);
        case 5:
            return launch<T, 5>(p, o.hd_variant, ctas, s);
        default:
            return fail(GTB_ERR_ARG, "gtb_hori_diff: hd.stages must be in 2..5");
        }
    }

} // namespace

GTB_API int gtb_hori_diff_f64(const gtb_field *in, const gtb_field *coeff, const gtb_field *out, int ni, int nj,
    int nk, void *stream) {
    return hori_diff<double>(in, coeff, out, ni, nj, nk, stream);
}

GTB_API int gtb_hori_diff_f32(const gtb_field *in, const gtb_field *coeff, const gtb_field *out, int ni, int nj,
    int nk, void *stream) {
    return hori_diff<float>(in, coeff, out, ni, nj, nk, stream);
}

// ------------------------------------------------------------------------------------ simple_hori_diff.cpp:25-61
// Two stages (wlap on the 1-extended domain, divflux), one ij_cached temporary, j-only coefficients crlato / crlatu.
// One launch: persistent 64x4-thread CTAs walk (tile, level) items; the halo-2 tile of `in` is staged in shared
// memory once, the laplacian tile (the reference's ij_cache) is computed into shared memory on the 1-extended tile,
// then every thread produces two outputs.  12 + 12 loads per point in the reference's kernel become one global load
// of `in`, one of `coeff` and one store.  Arithmetic in the functors' order (no FMA contraction): bit-identical to
// oracle/gt_oracle.c.
namespace {
    constexpr int SH_TI = 64, SH_TJ = 8, SH_TX = 64, SH_TY = 4;

    template <class T>
    struct shd_params {
        const T *in, *coeff, *crlato, *crlatu;
        T *out;
        int64_t in_sj, in_sk, co_sj, co_sk, out_sj, out_sk, cro_sj, cru_sj;
        int ni, nj, nk, tiles_i, tiles_j;
        int64_t items;
    };

    template <class T>
    __global__ void __launch_bounds__(SH_TX *SH_TY) shd_kernel(const shd_params<T> p) {
        constexpr int IW = SH_TI + 4 + 1, LW = SH_TI + 2 + 1; // padded row widths
        __shared__ T in_t[SH_TJ + 4][IW];
        __shared__ T lap_t[SH_TJ + 2][LW];
        const int tid = threadIdx.y * SH_TX + threadIdx.x;
        for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int ti = (int)(item % p.tiles_i);
            const int64_t r = item / p.tiles_i;
            const int tj = (int)(r % p.tiles_j), k = (int)(r / p.tiles_j);
            const int i0 = ti * SH_TI, j0 = tj * SH_TJ;
            const T *in_k = p.in + (int64_t)k * p.in_sk;
            for (int idx = tid; idx < (SH_TJ + 4) * (SH_TI + 4); idx += SH_TX * SH_TY) {
                const int jj = idx / (SH_TI + 4), ii = idx - jj * (SH_TI + 4);
                const int gi = i0 - 2 + ii, gj = j0 - 2 + jj;
                in_t[jj][ii] = gi < p.ni + 2 && gj < p.nj + 2 ? in_k[gi + (int64_t)gj * p.in_sj] : T(0);
            }
            __syncthreads();
            for (int idx = tid; idx < (SH_TJ + 2) * (SH_TI + 2); idx += SH_TX * SH_TY) { // wlap_function :25-41
                const int jj = idx / (SH_TI + 2), ii = idx - jj * (SH_TI + 2);
                const int gi = i0 - 1 + ii, gj = j0 - 1 + jj;
                if (gi <= p.ni && gj <= p.nj) {
                    const T c = in_t[jj + 1][ii + 1];
                    lap_t[jj][ii] = in_t[jj + 1][ii + 2] + in_t[jj + 1][ii] - T(2) * c +
                                    p.crlato[(int64_t)gj * p.cro_sj] * (in_t[jj + 2][ii + 1] - c) +
                                    p.crlatu[(int64_t)gj * p.cru_sj] * (in_t[jj][ii + 1] - c);
                }
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < SH_TJ / SH_TY; ++h) { // divflux_function :43-61
                const int jj = threadIdx.y + h * SH_TY, ii = threadIdx.x;
                const int gi = i0 + ii, gj = j0 + jj;
                if (gi < p.ni && gj < p.nj) {
                    const T c = lap_t[jj + 1][ii + 1];
                    const T cro = p.crlato[(int64_t)gj * p.cro_sj];
                    const T fluxx = lap_t[jj + 1][ii + 2] - c;
                    const T fluxx_m = c - lap_t[jj + 1][ii];
                    const T fluxy = cro * (lap_t[jj + 2][ii + 1] - c);
                    const T fluxy_m = cro * (c - lap_t[jj][ii + 1]);
                    p.out[gi + (int64_t)gj * p.out_sj + (int64_t)k * p.out_sk] =
                        in_t[jj + 2][ii + 2] +
                        ((fluxx_m - fluxx) + (fluxy_m - fluxy)) * p.coeff[gi + (int64_t)gj * p.co_sj + (int64_t)k * p.co_sk];
                }
            }
            __syncthreads(); // the tiles are rewritten by the next item
        }
    }

    template <class T>
    int simple_hori_diff(const gtb_field *in, const gtb_field *coeff, const gtb_field *crlato, const gtb_field *crlatu,
        const gtb_field *out, int ni, int nj, int nk, void *stream) {
        const gtb_field *all[5] = {in, coeff, crlato, crlatu, out};
        for (auto f : all)
            if (!field_ok(f))
                return fail(GTB_ERR_ARG, "gtb_simple_hori_diff: null field");
        if (in->stride_i != 1 || coeff->stride_i != 1 || out->stride_i != 1)
            return fail(GTB_ERR_LAYOUT, "gtb_simple_hori_diff: stride_i must be 1 (i is the unit-stride axis of storage::gpu)");
        if (ni < 0 || nj < 0 || nk < 0)
            return fail(GTB_ERR_ARG, "gtb_simple_hori_diff: negative size");
        if (out->ptr == in->ptr || out->ptr == coeff->ptr)
            return fail(GTB_ERR_ARG, "gtb_simple_hori_diff: out must not alias in / coeff");
        device_state *d = dev();
        if (!d)
            return GTB_ERR_CUDA;
        if (ni == 0 || nj == 0 || nk == 0)
            return GTB_OK;
        if (opts().hd_variant != 1) { // the TMA-staged register-tile kernel of hori_diff with the simple functors
            hd_params<T> q;
            q.in = static_cast<const T *>(in->ptr), q.coeff = static_cast<const T *>(coeff->ptr);
            q.out = static_cast<T *>(out->ptr);
            q.in_sj = in->stride_j, q.in_sk = in->stride_k, q.co_sj = coeff->stride_j, q.co_sk = coeff->stride_k;
            q.out_sj = out->stride_j, q.out_sk = out->stride_k;
            q.ni = ni, q.nj = nj, q.nk = nk;
            q.tiles_i = ceil_div(ni, BI), q.tiles_j = ceil_div(nj, BJ);
            q.crlato = static_cast<const T *>(crlato->ptr), q.crlatu = static_cast<const T *>(crlatu->ptr);
            q.cro_sj = crlato->stride_j, q.cru_sj = crlatu->stride_j;
            q.gate = stencil_gate();
            using L = layout<T>;
            constexpr int STAGES = 3;
            const int smem = STAGES * L::stage_bytes + 16 * STAGES;
            const int64_t items = (int64_t)q.tiles_i * q.tiles_j * nk;
            CUtensorMap map_in, map_co;
            if (items < ((int64_t)1 << 31) &&
                make_map<T>(&map_in, q.in, q.in_sj, q.in_sk, L::lead, 2, (int64_t)L::lead + ni + 2, nj + 4, nk, L::in_w, IN_H) &&
                make_map<T>(&map_co, q.coeff, q.co_sj, q.co_sk, 0, 0, ni, nj, nk, BI, BJ)) {
                int grid = stencil_sms(d) * 2;
                if (grid > items)
                    grid = (int)items;
                q.step_i = grid % q.tiles_i;
                q.step_j = (grid / q.tiles_i) % q.tiles_j;
                q.step_k = (grid / q.tiles_i) / q.tiles_j;
                auto kernel = hd_tma_kernel<T, STAGES, true>;
                int st = prepare_kernel(kernel, smem);
                if (st)
                    return st;
                GTB_CUDA(launch_pdl_if(false, kernel, dim3(grid), dim3(THREADS), (size_t)smem, as_stream(stream), map_in,
                    map_co, q)); // two CTAs per SM: see hori_diff
                count_launch();
                return check_launch("hd_tma_kernel (simple_hori_diff)");
            }
        }
        shd_params<T> p; // layouts TMA cannot address (or hd.variant = 1): plain shared-memory tile kernel
        p.in = static_cast<const T *>(in->ptr), p.coeff = static_cast<const T *>(coeff->ptr);
        p.crlato = static_cast<const T *>(crlato->ptr), p.crlatu = static_cast<const T *>(crlatu->ptr);
        p.out = static_cast<T *>(out->ptr);
        p.in_sj = in->stride_j, p.in_sk = in->stride_k, p.co_sj = coeff->stride_j, p.co_sk = coeff->stride_k;
        p.out_sj = out->stride_j, p.out_sk = out->stride_k;
        p.cro_sj = crlato->stride_j, p.cru_sj = crlatu->stride_j;
        p.ni = ni, p.nj = nj, p.nk = nk;
        p.tiles_i = ceil_div(ni, SH_TI), p.tiles_j = ceil_div(nj, SH_TJ);
        p.items = (int64_t)p.tiles_i * p.tiles_j * nk;
        int64_t grid = (int64_t)stencil_sms(d) * 8;
        if (grid > p.items)
            grid = p.items;
        shd_kernel<T><<<(unsigned)grid, dim3(SH_TX, SH_TY), 0, as_stream(stream)>>>(p);
        count_launch();
        return check_launch("shd_kernel");
    }
} // namespace

GTB_API int gtb_simple_hori_diff_f64(const gtb_field *in, const gtb_field *coeff, const gtb_field *crlato,
    const gtb_field *crlatu, const gtb_field *out, int ni, int nj, int nk, void *stream) {
    return simple_hori_diff<double>(in, coeff, crlato, crlatu, out, ni, nj, nk, stream);
}

GTB_API int gtb_simple_hori_diff_f32(const gtb_field *in, const gtb_field *coeff, const gtb_field *crlato,
    const gtb_field *crlatu, const gtb_field *out, int ni, int nj, int nk, void *stream) {
    return simple_hori_diff<float>(in, coeff, crlato, crlatu, out, ni, nj, nk, stream);
}
