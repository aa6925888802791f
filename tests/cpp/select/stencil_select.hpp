/*
 * stencil_select.hpp -- what a GridTools maintainer adds to tests/include/stencil_select.hpp for the B200 backend:
 * one more case next to GT_STENCIL_GPU.  This directory precedes the reference's
 * tests/include on the include path, so the reference's regression sources are compiled UNCHANGED.
 */
#pragma once

#if defined(GT_STENCIL_B200)
#ifndef GT_STORAGE_GPU
#define GT_STORAGE_GPU
#endif
#ifndef GT_TIMER_CUDA
#define GT_TIMER_CUDA
#endif
#include <gtb200/stencil/b200_select.hpp>
namespace {
    using stencil_backend_t = gridtools::stencil::b200<>;
}
#elif defined(GT_STENCIL_EMULATED)
// Test infrastructure, no GPU: the per-thread bodies of the fused generic path (b200_fused.hpp) run by emulated CTAs on
// the host (emulated_cta.hpp: one OpenMP team per CTA), behind the reference's unchanged regression sources.
#ifndef GT_STORAGE_CPU_IFIRST
#define GT_STORAGE_CPU_IFIRST
#endif
#ifndef GT_TIMER_OMP
#define GT_TIMER_OMP
#endif
#include <gridtools/common/timer/timer_omp.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include "../emulated_cta.hpp"
namespace emulated {
    template <class Geo>
    gridtools::storage::cpu_ifirst backend_storage_traits(backend<Geo>);
    template <class Geo>
    gridtools::timer_omp backend_timer_impl(backend<Geo>);
    template <class Geo>
    char const *backend_name(backend<Geo> const &) {
        return "b200_fused_on_emulated_ctas";
    }
    template <class Geo>
    std::false_type backend_supports_icosahedral(backend<Geo>);
    template <class Geo>
    std::true_type backend_supports_vertical_stencils(backend<Geo>);
} // namespace emulated
namespace {
    using stencil_backend_t = emulated::backend<>;
}
#endif

#include_next <stencil_select.hpp>
