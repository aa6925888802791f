/*
 * gtb200/stencil/b200.hpp -- the B200 backend tag for GridTools' `stencil::run` / `run_single_stage`.
 *
 *     #include <gridtools/stencil/cartesian.hpp>
 *     #include <gtb200/stencil/b200.hpp>
 *     gridtools::stencil::run(spec, gridtools::stencil::b200<>(), grid, fields...);
 *
 * A GridTools backend is a tag type plus one ADL-visible function template
 * `gridtools_backend_entry_point(tag, be_spec, grid, data_stores)` (stencil/core/backend.hpp:48, stencil/README.md);
 * this header provides exactly that and nothing above it: frontend, extent analysis, interval splitting and the
 * be_api (stencil/be_api.hpp) are consumed unchanged.  It replaces stencil/gpu/entry_point.hpp:254-260.
 *
 * Two execution paths behind the tag:
 *
 *  1. NAMED KERNELS.  If the ordered list of user functors of the spec has been registered with
 *     GTB200_REGISTER_SPEC(kernel, functors...), the whole spec (all stages, all multi-stages) is executed by ONE
 *     hand-written sm_100a kernel of libgtb200.so through the C ABI of include/gtb200.h (TMA-staged fused
 *     horizontal diffusion, TMA-streamed forward/backward vertical advection, ...).  The fields passed to `run`
 *     must be in the order of the corresponding C-ABI function, which is the order the reference's own specs use
 *     (horizontal_diffusion.cpp:98-106, horizontal_diffusion_fused.cpp:86-95, vertical_advection_dycore.cpp:140-149,
 *     tridiagonal.cpp:83-97, copy_stencil.cpp:42-47).  User functors are compile-time types a shared library cannot see, hence the one-line
 *     registration next to the functor definitions.
 *
 *  2. GENERIC PATHS (need nvcc: the user functors are instantiated inside a __global__ template in the user's
 *     translation unit).
 *     a. FUSED (b200_fused.hpp, the default where it applies): one launch per multi-stage, `__syncthreads` where
 *        be_api's need_sync asks for it, ij caches as shared-memory tiles, k caches as register windows with run-time
 *        checked fill / flush, forward / backward sweeps by one thread per column.
 *     b. STAGE BY STAGE: specs the fused path does not take (`fused::fusable`: non-cached temporaries read at IJ
 *        offsets, sweeps with IJ extents, k caches in parallel multi-stages) or `b200<..., gtb200::stage_by_stage>`:
 *        be_api::make_split_view, one launch per stage over the extent-extended IJ domain, k levels in parallel for
 *        execute_parallel and swept by one thread per column for execute_forward/backward, temporaries in device
 *        memory -- the semantics of the reference's `naive` backend (stencil/naive.hpp:32-78) executed on the GPU;
 *        ij/k caches are honoured as what they are semantically, plain temporaries.
 *
 * Errors: a non-zero status of the C ABI becomes the std::runtime_error the reference throws from GT_CUDA_CHECK
 * (common/cuda_util.hpp:20-35).  Launches go to the legacy default stream and do not synchronise, like the
 * reference (stencil/gpu/launch_kernel.hpp:161, common/cuda_util.hpp:79-96); `b200<stream_getter>` may supply
 * another stream.
 */
#pragma once

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

#include <gridtools/common/hymap.hpp>
#include <gridtools/common/integral_constant.hpp>
#include <gridtools/common/tuple_util.hpp>
#include <gridtools/meta.hpp>
#include <gridtools/sid/allocator.hpp>
#include <gridtools/sid/composite.hpp>
#include <gridtools/sid/concept.hpp>
#include <gridtools/sid/contiguous.hpp>
#include <gridtools/sid/sid_shift_origin.hpp>
#include <gridtools/stencil/be_api.hpp>
#include <gridtools/stencil/common/dim.hpp>
#include <gridtools/stencil/core/functor_metafunctions.hpp>
#include <gridtools/stencil/frontend/cartesian/stage.hpp>

#ifdef __CUDACC__
#include <gridtools/common/cuda_util.hpp>
#endif

#include "../../gtb200.h"
#include "b200_shapes.hpp"
#ifdef __CUDACC__
#include "b200_fused.hpp"
#endif

namespace gtb200 {

    /// Kernels of libgtb200.so a whole spec can be bound to.
    enum class kernel { none, copy, hori_diff, hori_diff_fused, simple_hori_diff, vert_adv, tridiagonal, prepare_tracers };

    /// Primary template: a spec made of these user functors (in stage order, duplicates removed) has no named kernel.
    template <class FunctorList>
    struct named_spec : std::integral_constant<kernel, kernel::none> {};

    struct default_stream {
        void *operator()() const { return nullptr; } // legacy default stream, like the reference
    };

    /// How specs without a named kernel are executed (second template argument of stencil::b200).
    struct fused_when_possible {};
    struct stage_by_stage {};
#ifdef __CUDACC__
    /// IJ block and levels per CTA of the fused generic path (third template argument of stencil::b200).
    template <int BI,
        int BJ,
        int KB,
        int SweepUnroll = 3,
        bool ChainSweeps = true,
        int Prefetch = 4,
        bool PrefetchL1 = true,
        int ParallelPrefetch = 0,
        bool StageReadOnly = true,
        bool RegisterTiles = true,
        bool L2Hints = false>
    using block_geometry = ::gridtools::stencil::b200_backend::fused::geometry<BI,
        BJ,
        KB,
        SweepUnroll,
        ChainSweeps,
        Prefetch,
        PrefetchL1,
        ParallelPrefetch,
        StageReadOnly,
        RegisterTiles,
        L2Hints>;
    using default_geometry = ::gridtools::stencil::b200_backend::fused::geometry<>;
#else
    struct default_geometry {};
#endif

    inline void check(int status, const char *what) {
        if (status != GTB_OK)
            throw std::runtime_error(std::string(what) + ": " + gtb_last_error());
    }
} // namespace gtb200

/// Binds a spec -- identified by its user functors in stage order -- to a named kernel.  Use at global scope, after
/// the functor definitions:
///     GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, lap_function, flx_function, fly_function, out_function)
#define GTB200_REGISTER_SPEC(KERNEL, ...)                                                     \
    template <>                                                                               \
    struct gtb200::named_spec<::gridtools::meta::list<__VA_ARGS__>>                           \
        : std::integral_constant<::gtb200::kernel, KERNEL> {}

namespace gridtools {
    namespace stencil {
        namespace b200_backend {
            namespace gt = ::gridtools;

            // ---------------------------------------------------------------- which user functors make up the spec
            template <class F>
            struct strip_bound {
                using type = F;
            };
            template <class F, class Param>
            struct strip_bound<core::bound_functor<F, Param>> {
                using type = F;
            };
            template <class Stage>
            struct stage_functor;
            template <class F, class PlhMap>
            struct stage_functor<cartesian::stage_impl_::stage<F, PlhMap>> {
                using type = typename strip_bound<F>::type;
            };
            template <class Stage>
            using stage_functor_t = typename stage_functor<Stage>::type;
            template <class Cell>
            using cell_functors = meta::transform<stage_functor_t, typename Cell::funs_t>;

            // Spec = list of multi-stage matrices; a matrix = list of stage rows; a row = list of cells (one per
            // elementary k interval).
            template <class Spec>
            using spec_functors = meta::dedup<
                meta::flatten<meta::transform<cell_functors, meta::flatten<meta::flatten<meta::rename<meta::list, Spec>>>>>>;

            // ---------------------------------------------------------------- raw views of the data stores
            template <class Sid>
            gtb_field as_field(Sid &sid_) {
                auto strides = sid::get_strides(sid_);
                gtb_field f;
                f.ptr = const_cast<void *>(static_cast<const void *>(sid::get_origin(sid_)()));
                f.stride_i = sid::get_stride<dim::i>(strides);
                f.stride_j = sid::get_stride<dim::j>(strides);
                f.stride_k = sid::get_stride<dim::k>(strides);
                return f;
            }
            template <class Sid>
            using element_of = std::remove_cv_t<std::remove_pointer_t<decltype(sid::get_origin(std::declval<Sid &>())())>>;

            using ::gtb200::kernel;
            template <kernel K>
            using kernel_c = std::integral_constant<kernel, K>;

            // copy_stencil.cpp:42-47 : run_single_stage(copy_functor(), backend, grid, in, out)
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::copy>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 2, "copy takes (in, out)");
                auto in = as_field(tuple_util::get<0>(ds)), out = as_field(tuple_util::get<1>(ds));
                using T = element_of<std::decay_t<decltype(tuple_util::get<1>(ds))>>;
                gtb200::check(
                    gtb_copy(&in, &out, grid.i_size(), grid.j_size(), grid.k_size(), (int)sizeof(T), stream), "gtb_copy");
            }

            // horizontal_diffusion.cpp:98-106 : run(spec, backend, grid, in, coeff, out)
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::hori_diff>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 3, "hori_diff takes (in, coeff, out)");
                auto in = as_field(tuple_util::get<0>(ds)), coeff = as_field(tuple_util::get<1>(ds)),
                     out = as_field(tuple_util::get<2>(ds));
                using T = element_of<std::decay_t<decltype(tuple_util::get<2>(ds))>>;
                static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "float or double");
                int st = std::is_same<T, double>::value
                             ? gtb_hori_diff_f64(&in, &coeff, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream)
                             : gtb_hori_diff_f32(&in, &coeff, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream);
                gtb200::check(st, "gtb_hori_diff");
            }

            // horizontal_diffusion_fused.cpp:86-95 : run_single_stage(out_function(), backend, grid, out, in, coeff) -- the
            // same stencil written as one stage that evaluates lap / flx / fly through call<>; same kernel, other
            // argument order
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::hori_diff_fused>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 3, "hori_diff_fused takes (out, in, coeff)");
                auto out = as_field(tuple_util::get<0>(ds)), in = as_field(tuple_util::get<1>(ds)),
                     coeff = as_field(tuple_util::get<2>(ds));
                using T = element_of<std::decay_t<decltype(tuple_util::get<0>(ds))>>;
                static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "float or double");
                int st = std::is_same<T, double>::value
                             ? gtb_hori_diff_f64(&in, &coeff, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream)
                             : gtb_hori_diff_f32(&in, &coeff, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream);
                gtb200::check(st, "gtb_hori_diff");
            }

            // simple_hori_diff.cpp:63-88 : run(spec, backend, grid, coeff, in, out, crlato, crlatu); crlato / crlatu are
            // j-only stores (builder selector<0,1,0>): their SIDs have no i / k stride
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::simple_hori_diff>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 5,
                    "simple_hori_diff takes (coeff, in, out, crlato, crlatu)");
                auto coeff = as_field(tuple_util::get<0>(ds)), in = as_field(tuple_util::get<1>(ds)),
                     out = as_field(tuple_util::get<2>(ds)), crlato = as_field(tuple_util::get<3>(ds)),
                     crlatu = as_field(tuple_util::get<4>(ds));
                using T = element_of<std::decay_t<decltype(tuple_util::get<2>(ds))>>;
                static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "float or double");
                int st = std::is_same<T, double>::value
                             ? gtb_simple_hori_diff_f64(
                                   &in, &coeff, &crlato, &crlatu, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream)
                             : gtb_simple_hori_diff_f32(
                                   &in, &coeff, &crlato, &crlatu, &out, grid.i_size(), grid.j_size(), grid.k_size(), stream);
                gtb200::check(st, "gtb_simple_hori_diff");
            }

            // vertical_advection_dycore.cpp:140-149 : run(spec, backend, grid, utens_stage, u_stage, wcon, u_pos, utens,
            // dtr_stage) with dtr_stage a global_parameter
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::vert_adv>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 6,
                    "vert_adv takes (utens_stage, u_stage, wcon, u_pos, utens, dtr_stage)");
                gtb_field f[5] = {as_field(tuple_util::get<0>(ds)),
                    as_field(tuple_util::get<1>(ds)),
                    as_field(tuple_util::get<2>(ds)),
                    as_field(tuple_util::get<3>(ds)),
                    as_field(tuple_util::get<4>(ds))};
                using T = element_of<std::decay_t<decltype(tuple_util::get<0>(ds))>>;
                static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "float or double");
                const T dtr = *sid::get_origin(tuple_util::get<5>(ds))();
                int st;
                if (std::is_same<T, double>::value)
                    st = gtb_vert_adv_f64(
                        &f[0], &f[1], &f[2], &f[3], &f[4], (double)dtr, grid.i_size(), grid.j_size(), grid.k_size(), stream);
                else
                    st = gtb_vert_adv_f32(
                        &f[0], &f[1], &f[2], &f[3], &f[4], (float)dtr, grid.i_size(), grid.j_size(), grid.k_size(), stream);
                gtb200::check(st, "gtb_vert_adv");
            }

            // tridiagonal.cpp:83-97 : run(spec, backend, grid, inf, diag, sup, rhs, out)
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::tridiagonal>, Grid const &grid, DataStores &ds, void *stream) {
                static_assert(tuple_util::size<DataStores>::value == 5, "tridiagonal takes (inf, diag, sup, rhs, out)");
                using T = element_of<std::decay_t<decltype(tuple_util::get<4>(ds))>>;
                static_assert(std::is_same<T, double>::value,
                    "gtb_tridiagonal_f64 is the only Thomas kernel (float specs take the generic path, see has_kernel)");
                gtb_field f[5] = {as_field(tuple_util::get<0>(ds)),
                    as_field(tuple_util::get<1>(ds)),
                    as_field(tuple_util::get<2>(ds)),
                    as_field(tuple_util::get<3>(ds)),
                    as_field(tuple_util::get<4>(ds))};
                gtb200::check(gtb_tridiagonal_f64(
                                  &f[0], &f[1], &f[2], &f[3], &f[4], grid.i_size(), grid.j_size(), grid.k_size(), stream),
                    "gtb_tridiagonal_f64");
            }

            // advection_pdbott_prepare_tracers.cpp:45-52 : one chunk of expandable_run<Factor>: (out_0..out_F-1, in_0..in_F-1, rho)
            template <class Grid, class DataStores>
            void run_named(kernel_c<kernel::prepare_tracers>, Grid const &grid, DataStores &ds, void *stream) {
                constexpr size_t F = (tuple_util::size<DataStores>::value - 1) / 2;
                gtb_field out[F], in[F], rho;
                size_t n = 0;
                tuple_util::for_each(
                    [&](auto &store) {
                        gtb_field f = as_field(store);
                        if (n < F)
                            out[n] = f;
                        else if (n < 2 * F)
                            in[n - F] = f;
                        else
                            rho = f;
                        ++n;
                    },
                    ds);
                gtb200::check(gtb_prepare_tracers_f64(
                                  out, in, (int)F, &rho, grid.i_size(), grid.j_size(), grid.k_size(), stream),
                    "gtb_prepare_tracers_f64");
            }

            // ---------------------------------------------------------------- is the binding valid for THIS spec?
            constexpr shape::id shape_of(kernel k) {
                return k == kernel::copy               ? shape::id::copy
                       : k == kernel::hori_diff        ? shape::id::hori_diff
                       : k == kernel::hori_diff_fused  ? shape::id::hori_diff_fused
                       : k == kernel::simple_hori_diff ? shape::id::simple_hori_diff
                       : k == kernel::vert_adv         ? shape::id::vert_adv
                       : k == kernel::tridiagonal      ? shape::id::tridiagonal
                       : k == kernel::prepare_tracers  ? shape::id::prepare_tracers
                                                       : shape::id::none;
            }
            // element types the kernel exists for
            template <kernel K, class T>
            using has_kernel = std::integral_constant<bool,
                K == kernel::copy ? (sizeof(T) == 4 || sizeof(T) == 8)
                : (K == kernel::tridiagonal || K == kernel::prepare_tracers)
                    ? std::is_same<T, double>::value
                    : (std::is_same<T, double>::value || std::is_same<T, float>::value)>;

            /// A spec runs on a named kernel iff its functor list is registered (GTB200_REGISTER_SPEC), the kernel exists
            /// for its element type, AND the spec is exactly the reference spec that kernel implements (b200_shapes.hpp):
            /// same wiring of placeholders to stage arguments, caches, intervals, extents and run() argument order.
            template <class Spec, class Grid, class DataStores>
            struct named_kernel_of {
                static constexpr kernel registered = ::gtb200::named_spec<spec_functors<Spec>>::value;
                using first_store_t = std::decay_t<decltype(tuple_util::get<0>(std::declval<DataStores &>()))>;
                using T = element_of<first_store_t>;
                template <kernel K>
                static constexpr bool usable() {
                    return has_kernel<K, T>::value &&
                           shape::matches<shape_of(K), spec_functors<Spec>, T, Spec, Grid, DataStores>::value;
                }
                static constexpr kernel value = registered != kernel::none && usable<registered>() ? registered : kernel::none;
            };

#ifdef __CUDACC__
            // ---------------------------------------------------------------- generic path (stage by stage)
            constexpr int generic_block_i = 32, generic_block_j = 8;

            // One launch = one stage on the extent-extended IJ domain.  Parallel stages: blockIdx.z is the level
            // (the cell that owns it is found by walking the intervals).  Sweeps: one thread per column walks all
            // intervals in execution order.
            template <class Stage, class PtrHolder, class Strides, class KSizes>
            __global__ void generic_stage_kernel(
                PtrHolder holder, Strides strides, KSizes k_sizes, int_t i_size, int_t j_size) {
                const int_t i = blockIdx.x * generic_block_i + threadIdx.x;
                const int_t j = blockIdx.y * generic_block_j + threadIdx.y;
                if (i >= i_size || j >= j_size)
                    return;
                auto ptr = holder();
                sid::shift(ptr, sid::get_stride<dim::i>(strides), i);
                sid::shift(ptr, sid::get_stride<dim::j>(strides), j);
                if (be_api::is_parallel<typename Stage::execution_t>::value) {
                    int_t k = blockIdx.z;
                    bool done = false;
                    tuple_util::device::for_each(
                        [&](auto cell, int_t size) {
                            if (done)
                                return;
                            if (k < size) {
                                sid::shift(ptr, sid::get_stride<dim::k>(strides), k);
                                cell(ptr, strides);
                                done = true;
                            } else {
                                k -= size;
                                sid::shift(ptr, sid::get_stride<dim::k>(strides), size);
                            }
                        },
                        Stage::cells(),
                        k_sizes);
                } else {
                    tuple_util::device::for_each(
                        [&](auto cell, int_t size) {
                            for (int_t k = 0; k < size; ++k) {
                                cell(ptr, strides);
                                cell.inc_k(ptr, strides);
                            }
                        },
                        Stage::cells(),
                        k_sizes);
                }
            }

            template <class Stage, class Grid, class DataStores>
            void launch_generic_stage(Stage, Grid const &grid, DataStores &data_stores, cudaStream_t stream) {
                using extent_t = typename Stage::extent_t;
                using plh_map_t = typename Stage::plh_map_t;
                using keys_t = meta::rename<sid::composite::keys, meta::transform<meta::first, plh_map_t>>;
                auto composite = tuple_util::convert_to<keys_t::template values>(tuple_util::transform(
                    [&](auto info) { return sid::add_const(info.is_const(), at_key<decltype(info.plh())>(data_stores)); },
                    Stage::plh_map()));
                using ptr_diff_t = sid::ptr_diff_type<decltype(composite)>;
                auto strides = sid::get_strides(composite);
                ptr_diff_t offset{};
                sid::shift(offset, sid::get_stride<dim::i>(strides), extent_t::minus(dim::i()));
                sid::shift(offset, sid::get_stride<dim::j>(strides), extent_t::minus(dim::j()));
                sid::shift(offset, sid::get_stride<dim::k>(strides), grid.k_start(Stage::interval(), Stage::execution()));
                auto k_sizes = tuple_util::transform([&](auto cell) { return grid.k_size(cell.interval()); }, Stage::cells());
                const int_t i_size = grid.i_size(extent_t()), j_size = grid.j_size(extent_t());
                int_t k_total = 0;
                tuple_util::for_each([&](int_t n) { k_total += n; }, k_sizes);
                if (i_size <= 0 || j_size <= 0 || k_total <= 0)
                    return;
                const bool parallel = be_api::is_parallel<typename Stage::execution_t>::value;
                dim3 block(generic_block_i, generic_block_j);
                dim3 blocks((i_size + generic_block_i - 1) / generic_block_i,
                    (j_size + generic_block_j - 1) / generic_block_j,
                    parallel ? k_total : 1);
                auto holder = sid::get_origin(composite) + offset;
                generic_stage_kernel<Stage><<<blocks, block, 0, stream>>>(holder, strides, k_sizes, i_size, j_size);
                GT_CUDA_CHECK(cudaGetLastError());
            }

            template <class Spec, class Grid, class DataStores>
            void run_generic(Spec, Grid const &grid, DataStores external, void *stream) {
                using stages_t = be_api::make_split_view<Spec>;
                using tmp_plh_map_t = be_api::remove_caches_from_plh_map<typename stages_t::tmp_plh_map_t>;
                auto alloc = fused::cuda_launcher{static_cast<cudaStream_t>(stream)}.allocator(); // per-stream free lists
                // temporaries: whole (extent-extended) domain in device memory, origin at the first compute point
                auto temporaries = be_api::make_data_stores(tmp_plh_map_t(), [&](auto info) {
                    auto extent = info.extent();
                    auto interval = stages_t::interval();
                    auto num_colors = info.num_colors();
                    auto offsets = hymap::keys<dim::i, dim::j, dim::k>::make_values(
                        -extent.minus(dim::i()), -extent.minus(dim::j()), -grid.k_start(interval) - extent.minus(dim::k()));
                    // first key = stride 1 (stride_util::make_strides_from_sizes): i fastest, like the fields
                    auto sizes = hymap::keys<dim::i, dim::j, dim::c, dim::k>::make_values(
                        grid.i_size(extent), grid.j_size(extent), num_colors, grid.k_size(interval, extent));
                    using stride_kind = meta::list<decltype(extent), decltype(num_colors)>;
                    return sid::shift_sid_origin(
                        sid::make_contiguous<decltype(info.data()), ptrdiff_t, stride_kind>(alloc, sizes), offsets);
                });
                auto data_stores = hymap::concat(std::move(external), std::move(temporaries));
                for_each<stages_t>(
                    [&](auto stage) { launch_generic_stage(stage, grid, data_stores, static_cast<cudaStream_t>(stream)); });
            }
#endif

            // ---------------------------------------------------------------- the tag
            template <class StreamGetter = ::gtb200::default_stream,
                class Generic = ::gtb200::fused_when_possible,
                class Geometry = ::gtb200::default_geometry>
            struct b200 {
                template <class Spec, class Grid, class DataStores>
                static void dispatch(std::true_type /*named*/, Spec, Grid const &grid, DataStores &data_stores) {
                    run_named(kernel_c<named_kernel_of<Spec, Grid, DataStores>::value>(), grid, data_stores,
                        StreamGetter()());
                }
                template <class Spec, class Grid, class DataStores>
                static void dispatch(std::false_type, Spec spec, Grid const &grid, DataStores &data_stores) {
#ifdef __CUDACC__
                    constexpr bool fuse =
                        std::is_same<Generic, ::gtb200::fused_when_possible>::value && fused::fusable<Spec>::value;
                    b200::generic(std::integral_constant<bool, fuse>(), spec, grid, data_stores);
#else
                    static_assert(sizeof(Spec) == 0,
                        "stencil::b200: this spec is not bound to a named kernel (no GTB200_REGISTER_SPEC for its functors, "
                        "or it does not have the shape of the reference spec the kernel implements, b200_shapes.hpp); the "
                        "generic path instantiates the user functors in a CUDA kernel and needs nvcc");
#endif
                }

#ifdef __CUDACC__
                template <class Spec, class Grid, class DataStores>
                static void generic(std::true_type /*fused*/, Spec spec, Grid const &grid, DataStores &data_stores) {
                    fused::cuda_launcher launcher{static_cast<cudaStream_t>(StreamGetter()())};
                    fused::run_fused_spec<Geometry>(launcher, spec, grid, std::move(data_stores));
                }
                template <class Spec, class Grid, class DataStores>
                static void generic(std::false_type, Spec spec, Grid const &grid, DataStores &data_stores) {
                    run_generic(spec, grid, std::move(data_stores), StreamGetter()());
                }
#endif

                template <class Spec, class Grid, class DataStores>
                friend void gridtools_backend_entry_point(b200, Spec spec, Grid const &grid, DataStores data_stores) {
                    constexpr bool named = named_kernel_of<Spec, Grid, DataStores>::value != ::gtb200::kernel::none;
                    // GTB200_TRACE_DISPATCH=1: one line per spec instantiation telling which path it takes
                    static const bool traced = [] {
                        if (std::getenv("GTB200_TRACE_DISPATCH"))
                            std::fprintf(stderr, "stencil::b200: spec -> %s (kernel id %d, registered id %d)\n",
                                named ? "named kernel" : "generic path", (int)named_kernel_of<Spec, Grid, DataStores>::value,
                                (int)named_kernel_of<Spec, Grid, DataStores>::registered);
                        return true;
                    }();
                    (void)traced;
                    b200::dispatch(std::integral_constant<bool, named>(), spec, grid, data_stores);
                }
            };
        } // namespace b200_backend
        using b200_backend::b200;
    } // namespace stencil
} // namespace gridtools
