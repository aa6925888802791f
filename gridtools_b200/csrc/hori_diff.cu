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
// does not satisfy TMA's 16-byte stride/alignment rules, and serves as an A/B baseline ("hd.variant" option).
//
// Arithmetic follows the functor bodies operation by operation (no FMA contraction: the file is compiled with
// -fmad=false), so results are bit-identical to oracle/gt_oracle.c compiled with -ffp-contract=off.
#include "common.cuh"

using namespace gtb;

namespace {

    constexpr int BI = 64;  // tile extent along i (one warp row = 32 consecutive i)
    constexpr int BJ = 16;  // tile extent along j
    constexpr int R = 4;    // outputs per thread (consecutive j)
    constexpr int THREADS = BI * (BJ / R);
    constexpr int IN_W = BI + 4;
    constexpr int IN_H = BJ + 4;

    template <class T>
    struct layout {
        static constexpr int in_bytes = IN_W * IN_H * (int)sizeof(T);
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
        int64_t items;
    };

    struct item_pos {
        int i0, j0, k;
    };

    template <class T>
    __device__ __forceinline__ item_pos decode(const hd_params<T> &p, int64_t w) {
        // i fastest, then j, then k: CTAs that run side by side work on neighbouring tiles of the same level, so
        // the halo rows they share are still in L2.
        int ti = (int)(w % p.tiles_i);
        int64_t r = w / p.tiles_i;
        int tj = (int)(r % p.tiles_j);
        int k = (int)(r / p.tiles_j);
        return {ti * BI, tj * BJ, k};
    }

    // The four stages for one thread: column i = tx, rows j0+ty*R .. +R-1, reading the staged tiles.
    template <class T>
    __device__ __forceinline__ void compute_item(
        const hd_params<T> &p, const T *__restrict__ sin, const T *__restrict__ sco, item_pos it, int tx, int ty) {
        const int jl = ty * R;
        const T *c = sin + (jl + 2) * IN_W + tx + 2; // in(i, j0+jl)
        T c0[R + 4], cp1[R + 2], cm1[R + 2], cp2[R], cm2[R];
#pragma unroll
        for (int d = 0; d < R + 4; ++d)
            c0[d] = c[(d - 2) * IN_W];
#pragma unroll
        for (int d = 0; d < R + 2; ++d) {
            cp1[d] = c[(d - 1) * IN_W + 1];
            cm1[d] = c[(d - 1) * IN_W - 1];
        }
#pragma unroll
        for (int d = 0; d < R; ++d) {
            cp2[d] = c[d * IN_W + 2];
            cm2[d] = c[d * IN_W - 2];
        }
        T co[R];
#pragma unroll
        for (int d = 0; d < R; ++d)
            co[d] = sco[(jl + d) * BI + tx];

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
        const int i = it.i0 + tx;
        const int j = it.j0 + jl;
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

    // ------------------------------------------------------------------------------------------ TMA variant
    template <class T, int STAGES>
    __global__ void __launch_bounds__(THREADS, 2) hd_tma_kernel(const __grid_constant__ CUtensorMap map_in,
        const __grid_constant__ CUtensorMap map_co, const hd_params<T> p, int pad_in, int pad_co) {
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
        __syncthreads();

        const int64_t first = blockIdx.x, step = gridDim.x;
        const int64_t n_my = first < p.items ? (p.items - first + step - 1) / step : 0;

        auto issue = [&](int s, int64_t w) {
            item_pos it = decode(p, w);
            unsigned char *base = smem + s * L::stage_bytes;
            ptx::mbar_expect_tx(&full[s], L::in_bytes + L::co_bytes);
            ptx::tma_load_3d(base, &map_in, &full[s], it.i0 + pad_in, it.j0, it.k);
            ptx::tma_load_3d(base + L::in_alloc, &map_co, &full[s], it.i0 + pad_co, it.j0, it.k);
        };

        if (tid == 0) {
            for (int s = 0; s < STAGES && s < n_my; ++s)
                issue(s, first + s * step);
        }
        for (int64_t n = 0; n < n_my; ++n) {
            const int s = (int)(n % STAGES);
            const uint32_t parity = (uint32_t)((n / STAGES) & 1);
            ptx::mbar_wait(&full[s], parity);
            const unsigned char *base = smem + s * L::stage_bytes;
            compute_item<T>(p,
                reinterpret_cast<const T *>(base),
                reinterpret_cast<const T *>(base + L::in_alloc),
                decode(p, first + n * step),
                tx,
                ty);
            __syncthreads(); // every thread has consumed stage s: hand it back to the TMA unit
            if (tid == 0 && n + STAGES < n_my)
                issue(s, first + (n + STAGES) * step);
        }
    }

    // ------------------------------------------------------------------------------- cp.async (LDGSTS) variant
    template <class T, int STAGES>
    __global__ void __launch_bounds__(THREADS, 2) hd_cpasync_kernel(const hd_params<T> p) {
        using L = layout<T>;
        extern __shared__ __align__(128) unsigned char smem[];
        const int tid = threadIdx.x;
        const int tx = tid % BI, ty = tid / BI;
        const int64_t first = blockIdx.x, step = gridDim.x;
        const int64_t n_my = first < p.items ? (p.items - first + step - 1) / step : 0;

        auto issue = [&](int s, int64_t w) {
            item_pos it = decode(p, w);
            T *sin = reinterpret_cast<T *>(smem + s * L::stage_bytes);
            T *sco = reinterpret_cast<T *>(smem + s * L::stage_bytes + L::in_alloc);
            const T *gin = p.in + (int64_t)it.k * p.in_sk;
            for (int e = tid; e < IN_W * IN_H; e += THREADS) {
                int r = e / IN_W, c = e - r * IN_W;
                int i = it.i0 + c - 2, j = it.j0 + r - 2;
                bool ok = i < p.ni + 2 && j < p.nj + 2; // i, j >= -2 always
                const T *src = ok ? gin + i + (int64_t)j * p.in_sj : p.in;
                ptx::cp_async<sizeof(T)>(sin + e, src, ok);
            }
            const T *gco = p.coeff + (int64_t)it.k * p.co_sk;
            for (int e = tid; e < BI * BJ; e += THREADS) {
                int r = e / BI, c = e - r * BI;
                int i = it.i0 + c, j = it.j0 + r;
                bool ok = i < p.ni && j < p.nj;
                const T *src = ok ? gco + i + (int64_t)j * p.co_sj : p.coeff;
                ptx::cp_async<sizeof(T)>(sco + e, src, ok);
            }
        };

        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < n_my)
                issue(s, first + s * step);
            ptx::cp_async_commit();
        }
        for (int64_t n = 0; n < n_my; ++n) {
            const int s = (int)(n % STAGES);
            ptx::cp_async_wait<STAGES - 2>(); // this thread's copies for item n have landed
            __syncthreads();                  // ... everybody's have, and stage (n-1)%STAGES is free again
            if (n + STAGES - 1 < n_my)
                issue((int)((n + STAGES - 1) % STAGES), first + (n + STAGES - 1) * step);
            ptx::cp_async_commit();
            const unsigned char *base = smem + s * L::stage_bytes;
            compute_item<T>(p,
                reinterpret_cast<const T *>(base),
                reinterpret_cast<const T *>(base + L::in_alloc),
                decode(p, first + n * step),
                tx,
                ty);
        }
        ptx::cp_async_wait<0>();
    }

    // ------------------------------------------------------------------------------------------ host side
    template <class T>
    CUtensorMapDataType tma_dtype();
    template <>
    CUtensorMapDataType tma_dtype<double>() {
        return CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    }
    template <>
    CUtensorMapDataType tma_dtype<float>() {
        return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    }

    // Tensor map over a field whose coordinate 0 sits `lead_i` (+pad for 16-byte alignment) elements before the
    // origin in i and `lead_j` rows before it in j.  Returns false when the layout is not TMA-addressable.
    template <class T>
    bool make_map(CUtensorMap *map, int *pad, const T *origin, int64_t sj, int64_t sk, int lead_i, int lead_j,
        int64_t len_i, int64_t len_j, int64_t len_k, int box_i, int box_j) {
        auto enc = tensor_map_encoder();
        if (!enc)
            return false;
        constexpr int es = sizeof(T);
        if ((sj * es) % 16 != 0 || (sk * es) % 16 != 0 || sj <= 0 || sk <= 0)
            return false;
        uintptr_t a = reinterpret_cast<uintptr_t>(origin - lead_i - (int64_t)lead_j * sj);
        int extra = (int)((a % 16) / es); // move the base down to the previous 16-byte boundary
        a -= (uintptr_t)extra * es;
        if (a % 16 != 0)
            return false;
        // keep the descriptor self-consistent: a row of the tensor must fit inside the row pitch
        if ((len_i + extra) > sj || len_j * sj > sk)
            return false;
        *pad = extra;
        cuuint64_t dims[3] = {(cuuint64_t)(len_i + extra), (cuuint64_t)len_j, (cuuint64_t)len_k};
        cuuint64_t strides[2] = {(cuuint64_t)(sj * es), (cuuint64_t)(sk * es)};
        // a single k level is addressed with a k stride that may be smaller than sj * len_j for padded layouts;
        // TMA only needs the strides to be multiples of 16 bytes
        cuuint32_t box[3] = {(cuuint32_t)box_i, (cuuint32_t)box_j, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(map,
            tma_dtype<T>(),
            3,
            reinterpret_cast<void *>(a),
            dims,
            strides,
            box,
            estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS;
    }

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
    int launch(const hd_params<T> &p, int variant, int ctas_per_sm, cudaStream_t stream) {
        using L = layout<T>;
        device_state *d = dev();
        const int smem = STAGES * L::stage_bytes + 8 * STAGES;
        int64_t grid = (int64_t)d->sm_count * ctas_per_sm;
        if (grid > p.items)
            grid = p.items;
        if (variant != 1) {
            CUtensorMap map_in, map_co;
            int pad_in = 0, pad_co = 0;
            bool ok = make_map<T>(&map_in, &pad_in, p.in, p.in_sj, p.in_sk, 2, 2, p.ni + 4, p.nj + 4, p.nk, IN_W, IN_H) &&
                      make_map<T>(&map_co, &pad_co, p.coeff, p.co_sj, p.co_sk, 0, 0, p.ni, p.nj, p.nk, BI, BJ);
            if (ok) {
                auto kernel = hd_tma_kernel<T, STAGES>;
                int st = prepare_kernel(kernel, smem);
                if (st)
                    return st;
                kernel<<<(unsigned)grid, THREADS, smem, stream>>>(map_in, map_co, p, pad_in, pad_co);
                count_launch();
                return check_launch("hd_tma_kernel");
            }
            if (variant == 2)
                return fail(GTB_ERR_LAYOUT,
                    "gtb_hori_diff: hd.variant=2 (TMA) needs stride_j/stride_k that are multiples of 16 bytes");
        }
        auto kernel = hd_cpasync_kernel<T, STAGES>;
        int st = prepare_kernel(kernel, smem);
        if (st)
            return st;
        kernel<<<(unsigned)grid, THREADS, smem, stream>>>(p);
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
        p.items = (int64_t)p.tiles_i * p.tiles_j * nk;
        const options &o = opts();
        int stages = o.hd_stages ? o.hd_stages : 4;
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
