// copy.cu -- pointwise kernels: copy (copy_stencil.cpp:24-36) and prepare_tracers
// (advection_pdbott_prepare_tracers.cpp:23-34).  Pure streaming: every thread issues all of its loads before the
// first store, rows are contiguous along i so a warp moves 512 B (16-byte vectors) per instruction.
#include "common.cuh"

using namespace gtb;

namespace {

    constexpr int kTx = 32;  // threads along i (vectors)
    constexpr int kTy = 8;   // thread rows
    constexpr int kRows = 4; // rows per thread

    struct fld {
        char *ptr;
        int64_t sj, sk; // byte strides (stride_i == element size on the vector path)
    };

    // V = 16-byte vector (or the scalar element type on the generic path); si = byte stride along i.
    template <class V>
    __global__ void __launch_bounds__(kTx *kTy) copy_kernel(fld in, fld out, int64_t in_si, int64_t out_si, int nvi,
        int nj, int nk) {
        const int iv = blockIdx.y * kTx + threadIdx.x;
        const int64_t nrows = (int64_t)nj * nk;
        const int64_t row0 = (int64_t)blockIdx.x * (kTy * kRows) + threadIdx.y;
        if (iv >= nvi)
            return;
        V v[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            int64_t row = row0 + r * kTy;
            if (row < nrows) {
                int j = (int)(row % nj), k = (int)(row / nj);
                v[r] = *reinterpret_cast<const V *>(in.ptr + iv * in_si + j * in.sj + k * in.sk);
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            int64_t row = row0 + r * kTy;
            if (row < nrows) {
                int j = (int)(row % nj), k = (int)(row / nj);
                *reinterpret_cast<V *>(out.ptr + iv * out_si + j * out.sj + k * out.sk) = v[r];
            }
        }
    }

    constexpr int kMaxTracers = 16;
    struct tracer_args {
        fld out[kMaxTracers];
        fld in[kMaxTracers];
        fld rho;
        int n;
    };

    template <class V, int N>
    __device__ __forceinline__ void tracers_body(const tracer_args &a, int64_t off_i, int j, int k) {
        V r = *reinterpret_cast<const V *>(a.rho.ptr + off_i + j * a.rho.sj + k * a.rho.sk);
        V v[N];
#pragma unroll
        for (int t = 0; t < N; ++t)
            if (t < a.n)
                v[t] = *reinterpret_cast<const V *>(a.in[t].ptr + off_i + j * a.in[t].sj + k * a.in[t].sk);
#pragma unroll
        for (int t = 0; t < N; ++t)
            if (t < a.n) {
                V o;
                if constexpr (sizeof(V) == 16) {
                    o.x = r.x * v[t].x; // rho * in, advection_pdbott_prepare_tracers.cpp:31
                    o.y = r.y * v[t].y;
                } else {
                    o = r * v[t];
                }
                *reinterpret_cast<V *>(a.out[t].ptr + off_i + j * a.out[t].sj + k * a.out[t].sk) = o;
            }
    }

    template <class V>
    __global__ void __launch_bounds__(kTx *kTy) tracers_kernel(const __grid_constant__ tracer_args a, int nvi, int nj,
        int nk) {
        const int iv = blockIdx.y * kTx + threadIdx.x;
        const int64_t nrows = (int64_t)nj * nk;
        const int64_t row = (int64_t)blockIdx.x * kTy + threadIdx.y;
        if (iv >= nvi || row >= nrows)
            return;
        int j = (int)(row % nj), k = (int)(row / nj);
        if (a.n <= 4)
            tracers_body<V, 4>(a, (int64_t)iv * sizeof(V), j, k);
        else if (a.n <= 8)
            tracers_body<V, 8>(a, (int64_t)iv * sizeof(V), j, k);
        else
            tracers_body<V, kMaxTracers>(a, (int64_t)iv * sizeof(V), j, k);
    }

    bool vec_ok(const gtb_field *f, int es, int ni) {
        const int vec = 16 / es;
        return f->stride_i == 1 && ni % vec == 0 && reinterpret_cast<uintptr_t>(f->ptr) % 16 == 0 &&
               f->stride_j % vec == 0 && f->stride_k % vec == 0;
    }

    fld make_fld(const gtb_field *f, int es) { return {static_cast<char *>(f->ptr), f->stride_j * es, f->stride_k * es}; }

} // namespace

GTB_API int gtb_copy(const gtb_field *in, const gtb_field *out, int ni, int nj, int nk, int elem_size, void *stream) {
    if (!field_ok(in) || !field_ok(out))
        return fail(GTB_ERR_ARG, "gtb_copy: null field");
    if (elem_size != 4 && elem_size != 8)
        return fail(GTB_ERR_ARG, "gtb_copy: elem_size %d not in {4,8}", elem_size);
    if (ni < 0 || nj < 0 || nk < 0)
        return fail(GTB_ERR_ARG, "gtb_copy: negative size");
    if (!dev())
        return GTB_ERR_CUDA;
    if (ni == 0 || nj == 0 || nk == 0)
        return GTB_OK;
    const int64_t nrows = (int64_t)nj * nk;
    dim3 block(kTx, kTy);
    fld fi = make_fld(in, elem_size), fo = make_fld(out, elem_size);
    if (opts().copy_vec && vec_ok(in, elem_size, ni) && vec_ok(out, elem_size, ni)) {
        const int nvi = ni / (16 / elem_size);
        dim3 grid((unsigned)ceil_div((int)((nrows + kTy * kRows - 1) / (kTy * kRows)), 1), ceil_div(nvi, kTx));
        copy_kernel<uint4><<<grid, block, 0, as_stream(stream)>>>(fi, fo, 16, 16, nvi, nj, nk);
    } else {
        dim3 grid((unsigned)((nrows + kTy * kRows - 1) / (kTy * kRows)), ceil_div(ni, kTx));
        if (elem_size == 8)
            copy_kernel<uint64_t><<<grid, block, 0, as_stream(stream)>>>(
                fi, fo, in->stride_i * 8, out->stride_i * 8, ni, nj, nk);
        else
            copy_kernel<uint32_t><<<grid, block, 0, as_stream(stream)>>>(
                fi, fo, in->stride_i * 4, out->stride_i * 4, ni, nj, nk);
    }
    count_launch();
    return check_launch("gtb_copy");
}

GTB_API int gtb_prepare_tracers_f64(const gtb_field *out, const gtb_field *in, int n_tracers, const gtb_field *rho,
    int ni, int nj, int nk, void *stream) {
    if (!out || !in || !field_ok(rho) || n_tracers < 0)
        return fail(GTB_ERR_ARG, "gtb_prepare_tracers_f64: bad arguments");
    if (ni < 0 || nj < 0 || nk < 0)
        return fail(GTB_ERR_ARG, "gtb_prepare_tracers_f64: negative size");
    for (int t = 0; t < n_tracers; ++t) {
        if (!field_ok(&out[t]) || !field_ok(&in[t]))
            return fail(GTB_ERR_ARG, "gtb_prepare_tracers_f64: null field %d", t);
        if (out[t].stride_i != 1 || in[t].stride_i != 1)
            return fail(GTB_ERR_LAYOUT, "gtb_prepare_tracers_f64: stride_i must be 1");
    }
    if (rho->stride_i != 1)
        return fail(GTB_ERR_LAYOUT, "gtb_prepare_tracers_f64: stride_i must be 1");
    if (!dev())
        return GTB_ERR_CUDA;
    if (ni == 0 || nj == 0 || nk == 0 || n_tracers == 0)
        return GTB_OK;
    const int64_t nrows = (int64_t)nj * nk;
    dim3 block(kTx, kTy);
    for (int t0 = 0; t0 < n_tracers; t0 += kMaxTracers) {
        tracer_args a;
        a.n = n_tracers - t0 < kMaxTracers ? n_tracers - t0 : kMaxTracers;
        bool vec = vec_ok(rho, 8, ni);
        for (int t = 0; t < a.n; ++t) {
            a.out[t] = make_fld(&out[t0 + t], 8);
            a.in[t] = make_fld(&in[t0 + t], 8);
            vec = vec && vec_ok(&out[t0 + t], 8, ni) && vec_ok(&in[t0 + t], 8, ni);
        }
        a.rho = make_fld(rho, 8);
        if (vec) {
            dim3 grid((unsigned)((nrows + kTy - 1) / kTy), ceil_div(ni / 2, kTx));
            tracers_kernel<double2><<<grid, block, 0, as_stream(stream)>>>(a, ni / 2, nj, nk);
        } else {
            dim3 grid((unsigned)((nrows + kTy - 1) / kTy), ceil_div(ni, kTx));
            tracers_kernel<double><<<grid, block, 0, as_stream(stream)>>>(a, ni, nj, nk);
        }
        count_launch();
        int st = check_launch("gtb_prepare_tracers_f64");
        if (st)
            return st;
    }
    return GTB_OK;
}
