/*
 * gcl_select.hpp -- what a GridTools maintainer adds to tests/include/gcl_select.hpp for the B200 arch: one more
 * case next to GT_GCL_GPU / GT_GCL_CPU.  This directory precedes the reference's
 * tests/include on the include path, so the reference's test sources are compiled UNCHANGED; everything else comes
 * from the reference's own header through #include_next.
 */
#pragma once

#if defined(GT_GCL_B200)
#ifndef GT_STORAGE_GPU
#define GT_STORAGE_GPU
#endif
#ifndef GT_TIMER_CUDA
#define GT_TIMER_CUDA
#endif
#include <gridtools/gcl/halo_exchange.hpp>
#include <gtb200/boundaries/b200.hpp>
#include <gtb200/gcl/b200.hpp>
namespace {
    using gcl_arch_t = gridtools::gcl::b200;
}
#endif

#include_next <gcl_select.hpp>

#if defined(GT_GCL_B200)
namespace gridtools {
    namespace gcl {
        storage::gpu backend_storage_traits(b200 const &);
        timer_cuda backend_timer_impl(b200 const &);
        inline char const *backend_name(b200 const &) { return "b200"; }
    } // namespace gcl
} // namespace gridtools
#endif
