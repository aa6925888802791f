// functors.hpp -- user-side stencil functors for the C++ boundary tests.  In GridTools terms this is USER code: it
// restates the functors of the reference's regression suite (tests/regression/horizontal_diffusion.cpp:35-91,
// vertical_advection_dycore.cpp:32-122, tridiagonal.cpp:39-74, copy_stencil.cpp:24-36, simple_hori_diff.cpp:25-61)
// with identical operand order.  Each functor family exists twice: the plain one is registered with
// GTB200_REGISTER_SPEC and runs on the named sm_100a kernels, the `g_` twin is not registered and therefore goes
// through the generic stage-by-stage path of stencil::b200.
#pragma once

#include <gridtools/stencil/cartesian.hpp>

namespace user {
    namespace st = gridtools::stencil;
    using namespace gridtools::stencil;
    using namespace gridtools::stencil::cartesian;

    template <int Tag>
    struct copy_f {
        using in = in_accessor<0>;
        using out = inout_accessor<1>;
        using param_list = make_param_list<in, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(in());
        }
    };

    template <int Tag>
    struct lap_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, -1, 1>>;
        using param_list = make_param_list<out, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            using float_t = std::decay_t<decltype(eval(out()))>;
            eval(out()) = float_t{4} * eval(in()) - (eval(in(1, 0)) + eval(in(0, 1)) + eval(in(-1, 0)) + eval(in(0, -1)));
        }
    };
    template <int Tag>
    struct flx_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<0, 1, 0, 0>>;
        using lap = in_accessor<2, extent<0, 1, 0, 0>>;
        using param_list = make_param_list<out, in, lap>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto res = eval(lap(1, 0)) - eval(lap(0, 0));
            eval(out()) = res * (eval(in(1, 0)) - eval(in(0, 0))) > 0 ? 0 : res;
        }
    };
    template <int Tag>
    struct fly_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<0, 0, 0, 1>>;
        using lap = in_accessor<2, extent<0, 0, 0, 1>>;
        using param_list = make_param_list<out, in, lap>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto res = eval(lap(0, 1)) - eval(lap(0, 0));
            eval(out()) = res * (eval(in(0, 1)) - eval(in(0, 0))) > 0 ? 0 : res;
        }
    };
    template <int Tag>
    struct out_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1>;
        using flx = in_accessor<2, extent<-1, 0, 0, 0>>;
        using fly = in_accessor<3, extent<0, 0, -1, 0>>;
        using coeff = in_accessor<4>;
        using param_list = make_param_list<out, in, flx, fly, coeff>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(in()) - eval(coeff()) * (eval(flx()) - eval(flx(-1, 0)) + eval(fly()) - eval(fly(0, -1)));
        }
    };

    template <class T, int Tag>
    auto hori_diff_spec() {
        return [](auto in, auto coeff, auto out) {
            GT_DECLARE_TMP(T, lap, flx, fly);
            return st::execute_parallel()
                .ij_cached(lap, flx, fly)
                .stage(lap_f<Tag>(), lap, in)
                .stage(flx_f<Tag>(), flx, in, lap)
                .stage(fly_f<Tag>(), fly, in, lap)
                .stage(out_f<Tag>(), out, in, flx, fly, coeff);
        };
    }

    // ---- the same stencil as ONE stage, intermediate results through call<> (horizontal_diffusion_fused.cpp:23-84)
    template <int Tag>
    struct fused_flx_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 2, -1, 1>>;
        using param_list = make_param_list<out, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto hi = call<lap_f<Tag>>::with(eval, in(1, 0));
            auto lo = call<lap_f<Tag>>::with(eval, in(0, 0));
            auto flx = hi - lo;
            eval(out()) = flx * (eval(in(1, 0)) - eval(in(0, 0))) > 0 ? 0 : flx;
        }
    };
    template <int Tag>
    struct fused_fly_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, -1, 2>>;
        using param_list = make_param_list<out, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto hi = call<lap_f<Tag>>::with(eval, in(0, 1));
            auto lo = call<lap_f<Tag>>::with(eval, in(0, 0));
            auto fly = hi - lo;
            eval(out()) = fly * (eval(in(0, 1)) - eval(in(0, 0))) > 0 ? 0 : fly;
        }
    };
    template <int Tag>
    struct fused_out_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-2, 2, -2, 2>>;
        using coeff = in_accessor<2>;
        using param_list = make_param_list<out, in, coeff>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto flx_hi = call<fused_flx_f<Tag>>::with(eval, in(0, 0));
            auto flx_lo = call<fused_flx_f<Tag>>::with(eval, in(-1, 0));
            auto fly_hi = call<fused_fly_f<Tag>>::with(eval, in(0, 0));
            auto fly_lo = call<fused_fly_f<Tag>>::with(eval, in(0, -1));
            eval(out()) = eval(in()) - eval(coeff()) * (flx_hi - flx_lo + fly_hi - fly_lo);
        }
    };

    // ---- simple horizontal diffusion (simple_hori_diff.cpp:25-61): j-only coefficient fields
    template <int Tag>
    struct wlap_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, -1, 1>>;
        using crlato = in_accessor<2>;
        using crlatu = in_accessor<3>;
        using param_list = make_param_list<out, in, crlato, crlatu>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            using float_t = std::decay_t<decltype(eval(out()))>;
            eval(out()) = eval(in(1, 0)) + eval(in(-1, 0)) - float_t{2} * eval(in()) +
                          eval(crlato()) * (eval(in(0, 1)) - eval(in())) + eval(crlatu()) * (eval(in(0, -1)) - eval(in()));
        }
    };
    template <int Tag>
    struct divflux_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1>;
        using lap = in_accessor<2, extent<-1, 1, -1, 1>>;
        using crlato = in_accessor<3>;
        using coeff = in_accessor<4>;
        using param_list = make_param_list<out, in, lap, crlato, coeff>;
        template <class E>
        GT_FUNCTION static void apply(E &eval) {
            auto fluxx = eval(lap(1, 0)) - eval(lap());
            auto fluxx_m = eval(lap()) - eval(lap(-1, 0));
            auto fluxy = eval(crlato()) * (eval(lap(0, 1)) - eval(lap()));
            auto fluxy_m = eval(crlato()) * (eval(lap()) - eval(lap(0, -1)));
            eval(out()) = eval(in()) + ((fluxx_m - fluxx) + (fluxy_m - fluxy)) * eval(coeff());
        }
    };
    template <class T, int Tag>
    auto simple_hori_diff_spec() {
        return [](auto coeff, auto in, auto out, auto crlato, auto crlatu) {
            GT_DECLARE_TMP(T, lap);
            return st::execute_parallel()
                .ij_cached(lap)
                .stage(wlap_f<Tag>(), lap, in, crlato, crlatu)
                .stage(divflux_f<Tag>(), out, in, lap, crlato, coeff);
        };
    }

    // ---- vertical advection (vertical_advection_dycore.cpp), BET_M = BET_P = 0.5 (vertical_advection_defs.hpp)
    using va_axis_t = st::axis<1, st::axis_config::offset_limit<3>>;
    using va_full_t = va_axis_t::full_interval;

    template <int Tag>
    struct va_forward_f {
        using utens_stage = in_accessor<0>;
        using wcon = in_accessor<1, extent<0, 1, 0, 0, 0, 1>>;
        using u_stage = in_accessor<2, extent<0, 0, 0, 0, -1, 1>>;
        using u_pos = in_accessor<3>;
        using utens = in_accessor<4>;
        using dtr_stage = in_accessor<5>;
        using ccol = inout_accessor<6, extent<0, 0, 0, 0, -1, 0>>;
        using dcol = inout_accessor<7, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<utens_stage, wcon, u_stage, u_pos, utens, dtr_stage, ccol, dcol>;

        template <class E>
        GT_FUNCTION static auto dcol_base(E &&eval) {
            return eval(dtr_stage()) * eval(u_pos()) + eval(utens()) + eval(utens_stage());
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::first_level) {
            using T = std::decay_t<decltype(eval(ccol()))>;
            auto gcv = T(.25) * (eval(wcon(1, 0, 1)) + eval(wcon(0, 0, 1)));
            auto cs = gcv * T(.5);
            auto c = gcv * T(.5);
            auto b = eval(dtr_stage()) - c;
            auto correction = -cs * (eval(u_stage(0, 0, 1)) - eval(u_stage()));
            auto d = dcol_base(eval) + correction;
            auto divided = T(1) / b;
            eval(ccol()) = c * divided;
            eval(dcol()) = d * divided;
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::modify<1, -1>) {
            using T = std::decay_t<decltype(eval(ccol()))>;
            auto gav = -T(.25) * (eval(wcon(1, 0, 0)) + eval(wcon(0, 0, 0)));
            auto gcv = T(.25) * (eval(wcon(1, 0, 1)) + eval(wcon(0, 0, 1)));
            auto as = gav * T(.5);
            auto cs = gcv * T(.5);
            auto a = gav * T(.5);
            auto c = gcv * T(.5);
            auto b = eval(dtr_stage()) - a - c;
            auto correction = -as * (eval(u_stage(0, 0, -1)) - eval(u_stage())) -
                              cs * (eval(u_stage(0, 0, 1)) - eval(u_stage()));
            auto d = dcol_base(eval) + correction;
            auto divided = T(1) / (b - eval(ccol(0, 0, -1)) * a);
            eval(ccol()) = c * divided;
            eval(dcol()) = (d - eval(dcol(0, 0, -1)) * a) * divided;
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::last_level) {
            using T = std::decay_t<decltype(eval(ccol()))>;
            auto gav = -T(.25) * (eval(wcon(1, 0, 0)) + eval(wcon()));
            auto as = gav * T(.5);
            auto a = gav * T(.5);
            auto b = eval(dtr_stage()) - a;
            auto correction = -as * (eval(u_stage(0, 0, -1)) - eval(u_stage()));
            auto d = dcol_base(eval) + correction;
            auto divided = T(1) / (b - eval(ccol(0, 0, -1)) * a);
            eval(dcol()) = (d - eval(dcol(0, 0, -1)) * a) * divided;
        }
    };
    template <int Tag>
    struct va_backward_f {
        using utens_stage = inout_accessor<0>;
        using u_pos = in_accessor<1>;
        using dtr_stage = in_accessor<2>;
        using ccol = in_accessor<3>;
        using dcol = in_accessor<4>;
        using data_col = inout_accessor<5, extent<0, 0, 0, 0, 0, 1>>;
        using param_list = make_param_list<utens_stage, u_pos, dtr_stage, ccol, dcol, data_col>;
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::modify<0, -1>) {
            auto data = eval(dcol()) - eval(ccol()) * eval(data_col(0, 0, 1));
            eval(utens_stage()) = eval(dtr_stage()) * (data - eval(u_pos()));
            eval(data_col()) = data;
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::last_level) {
            eval(utens_stage()) = eval(dtr_stage()) * (eval(dcol()) - eval(u_pos()));
            eval(data_col()) = eval(dcol());
        }
    };

    template <class T, int Tag>
    auto vert_adv_spec() {
        return [](auto utens_stage, auto u_stage, auto wcon, auto u_pos, auto utens, auto dtr_stage) {
            GT_DECLARE_TMP(T, ccol, dcol, data_col);
            return st::multi_pass(
                st::execute_forward()
                    .k_cached(st::cache_io_policy::flush(), ccol, dcol)
                    .k_cached(st::cache_io_policy::fill(), u_stage)
                    .stage(va_forward_f<Tag>(), utens_stage, wcon, u_stage, u_pos, utens, dtr_stage, ccol, dcol),
                st::execute_backward().k_cached(data_col).stage(
                    va_backward_f<Tag>(), utens_stage, u_pos, dtr_stage, ccol, dcol, data_col));
        };
    }

    // ---- Thomas solve (tridiagonal.cpp)
    using td_axis_t = st::axis<1>;
    using td_full_t = td_axis_t::full_interval;
    template <int Tag>
    struct td_forward_f {
        using inf = in_accessor<0>;
        using diag = in_accessor<1>;
        using sup = inout_accessor<2, extent<0, 0, 0, 0, -1, 0>>;
        using rhs = inout_accessor<3, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<inf, diag, sup, rhs>;
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::modify<1, 0>) {
            eval(sup()) = eval(sup()) / (eval(diag()) - eval(sup(0, 0, -1)) * eval(inf()));
            eval(rhs()) =
                (eval(rhs()) - eval(inf()) * eval(rhs(0, 0, -1))) / (eval(diag()) - eval(sup(0, 0, -1)) * eval(inf()));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::first_level) {
            eval(sup()) = eval(sup()) / eval(diag());
            eval(rhs()) = eval(rhs()) / eval(diag());
        }
    };
    template <int Tag>
    struct td_backward_f {
        using out = inout_accessor<0, extent<0, 0, 0, 0, 0, 1>>;
        using sup = in_accessor<1>;
        using rhs = in_accessor<2>;
        using param_list = make_param_list<out, sup, rhs>;
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::modify<0, -1>) {
            eval(out()) = eval(rhs()) - eval(sup()) * eval(out(0, 0, 1));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::last_level) {
            eval(out()) = eval(rhs());
        }
    };
    template <int Tag>
    auto tridiagonal_spec() {
        return [](auto inf, auto diag, auto sup, auto rhs, auto out) {
            return st::multi_pass(st::execute_forward().stage(td_forward_f<Tag>(), inf, diag, sup, rhs),
                st::execute_backward().stage(td_backward_f<Tag>(), out, sup, rhs));
        };
    }
} // namespace user
