/*
 * gtb200/stencil/b200_select.hpp -- what the reference's test harness asks of a backend tag besides the entry point
 * (tests/include/stencil_select.hpp:129-231, tests/include/test_environment.hpp:82-117): the storage traits its
 * stores are built with, the timer `perftests` uses, a name for the report, and the capability probes.  Found by
 * ADL on the tag, exactly like the blocks for `gpu<>` (stencil_select.hpp:197-211) and `gpu_horizontal<>` (:213-231).
 *
 * A maintainer adds to tests/include/stencil_select.hpp
 *
 *     #elif defined(GT_STENCIL_B200)
 *     #ifndef GT_STORAGE_GPU
 *     #define GT_STORAGE_GPU
 *     #endif
 *     #ifndef GT_TIMER_CUDA
 *     #define GT_TIMER_CUDA
 *     #endif
 *     #include <gtb200/stencil/b200_select.hpp>
 *     namespace { using stencil_backend_t = gridtools::stencil::b200<>; }
 *
 * and the regression suite compiles against the new tag unchanged (plus one GTB200_REGISTER_SPEC line per spec that
 * should run on its hand-written kernel; every other spec takes the generic paths).
 */
#pragma once

#include <type_traits>

#include <gridtools/storage/gpu.hpp>

#include "b200.hpp"

namespace gridtools {
    class timer_cuda; // gridtools/common/timer/timer_cuda.hpp; the harness only needs the type (timer_select.hpp:13)

    namespace stencil {
        namespace b200_backend {
            // stores in the layout of storage::gpu (i stride 1, rows padded to 128 bytes): what the named kernels'
            // TMA boxes and the generic paths' coalesced rows assume (storage/gpu.hpp:39,78)
            template <class S, class G, class Geo>
            storage::gpu backend_storage_traits(b200<S, G, Geo>);

            // launches are asynchronous on a CUDA stream, so wall-clock timers would measure nothing
            template <class S, class G, class Geo>
            timer_cuda backend_timer_impl(b200<S, G, Geo>);

            template <class S, class G, class Geo>
            char const *backend_name(b200<S, G, Geo> const &) {
                return "b200";
            }

            // Cartesian grids only (icosahedral is out of scope); vertical stencils are supported
            template <class S, class G, class Geo>
            std::false_type backend_supports_icosahedral(b200<S, G, Geo>);
            template <class S, class G, class Geo>
            std::true_type backend_supports_vertical_stencils(b200<S, G, Geo>);
        } // namespace b200_backend
    } // namespace stencil
} // namespace gridtools
