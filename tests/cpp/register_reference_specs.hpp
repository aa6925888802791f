/*
 * register_reference_specs.hpp -- force-included (-include) in front of the reference's UNCHANGED regression sources to
 * bind their specs to the hand-written kernels: the one line per spec a maintainer adds after the functor definitions
 * (INTEGRATION.md section 1), here placed BEFORE them with forward declarations so that the sources stay untouched.
 * The functors of tests/regression/*.cpp live in the translation unit's anonymous namespace.
 */
#pragma once
#include <gtb200/stencil/b200.hpp>

namespace {
    struct copy_functor;                                                   // copy_stencil.cpp:23
    struct lap_function; struct flx_function; struct fly_function; struct out_function; // horizontal_diffusion*.cpp
    struct wlap_function; struct divflux_function;                         // simple_hori_diff.cpp:25,42
    class u_forward_function; class u_backward_function;                   // vertical_advection_dycore.cpp:32,96
    struct forward_thomas; struct backward_thomas;                         // tridiagonal.cpp:39,59
    struct prepare_tracers;                                                // advection_pdbott_prepare_tracers.cpp:23
} // namespace
GTB200_REGISTER_SPEC(gtb200::kernel::copy, copy_functor);
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, lap_function, flx_function, fly_function, out_function);
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff_fused, out_function);
GTB200_REGISTER_SPEC(gtb200::kernel::simple_hori_diff, wlap_function, divflux_function);
GTB200_REGISTER_SPEC(gtb200::kernel::vert_adv, u_forward_function, u_backward_function);
GTB200_REGISTER_SPEC(gtb200::kernel::tridiagonal, forward_thomas, backward_thomas);
GTB200_REGISTER_SPEC(gtb200::kernel::prepare_tracers, prepare_tracers);
