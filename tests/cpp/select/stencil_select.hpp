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
#endif

#include_next <stencil_select.hpp>
