// cases.hpp -- stencil cases for the generic paths of stencil::b200, written once against (storage traits, backend) so
// that the same case runs (a) on the device through st::b200<> (b200_regression.cu), (b) on the host through the
// emulated CTAs of the fused path (fused_emulation.cpp) and (c) on the reference's cpu_ifirst backend, which is what
// (a) and (b) are compared with.  The specs are those of functors.hpp (tag 1 = not bound to a named kernel) plus the
// k-cache patterns of the reference's unit tests (tests/unit_tests/stencil/frontend/cartesian/test_kcache_fill.cpp,
// test_kcache_flush.cpp, test_kcache_fill_and_flush.cpp, test_kcache_local.cpp: shifted sums through filled, flushed
// and local register windows, forward and backward) and two mixed specs that exercise shared-memory tiles beside
// plain temporaries and several elementary k intervals inside one parallel multi-stage.
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/sid/rename_dimensions.hpp>
#include <gridtools/stencil/global_parameter.hpp>
#include <gridtools/stencil/positional.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/sid.hpp>

#include "functors.hpp"

namespace cases {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;
    using namespace gridtools::stencil;
    using namespace gridtools::stencil::cartesian;
    using fun_t = std::function<double(int, int, int)>;

    template <class Traits, class T>
    auto make_store(int d0, int d1, int d2, int halo, fun_t f) {
        return gt::storage::builder<Traits>.template type<T>().dimensions(d0, d1, d2).halos(halo, halo, 0)
            .initializer([f](int i, int j, int k) { return T(f(i, j, k)); })
            .build();
    }

    inline auto ij_halos(int d0, int d1, int halo) {
        auto h = [&](int d) { return gt::halo_descriptor(halo, halo, halo, d - halo - 1, d); };
        return std::make_pair(h(d0), h(d1));
    }

    // tests/include/verifier.hpp:26-52 on the whole storage (halo included: a backend must not touch it)
    template <class A, class B>
    bool same(const char *name, A const &a, B const &b, int d0, int d1, int d2, double tol, int &failed) {
        auto va = a->const_host_view();
        auto vb = b->const_host_view();
        double worst = 0;
        long bad = 0;
        for (int k = 0; k < d2; ++k)
            for (int j = 0; j < d1; ++j)
                for (int i = 0; i < d0; ++i) {
                    double x = va(i, j, k), y = vb(i, j, k);
                    double d = std::fabs(x - y), s = std::fmax(std::fabs(x), std::fabs(y));
                    double rel = s > 0 ? d / s : 0;
                    if (tol == 0 ? x != y : !(d < tol || rel < tol))
                        ++bad;
                    if (rel > worst)
                        worst = rel;
                }
        std::printf("%-58s %s (mismatches %ld, worst rel %.3g)\n", name, bad ? "FAILED" : "ok", bad, worst);
        if (bad)
            ++failed;
        return bad == 0;
    }

    inline double ramp(int i, int j, int k) { return 1 + i + 0.5 * j + 0.25 * k * k; }

    // ------------------------------------------------------------------ specs of functors.hpp
    template <class T, class Traits, class Backend>
    auto hori_diff(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int k) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4. + 0.02 * k;
        };
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<Traits, T const>(d0, d1, nk, H, in_f);
        auto co = make_store<Traits, T const>(d0, d1, nk, H, [](int i, int j, int) { return 0.025 + 1e-4 * ((i + j) % 5); });
        auto out = make_store<Traits, T>(d0, d1, nk, H, [](int, int, int) { return -1.; });
        st::run(user::hori_diff_spec<T, 1>(), backend, grid, in, co, out);
        return out;
    }

    // horizontal_diffusion_fused.cpp:86-95: one stage, lap / flx / fly evaluated through call<>; must give what the
    // four-stage spec gives
    template <class T, int Tag = 1, class Traits, class Backend>
    auto hori_diff_fused(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int k) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4. + 0.02 * k;
        };
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<Traits, T const>(d0, d1, nk, H, in_f);
        auto co = make_store<Traits, T const>(d0, d1, nk, H, [](int i, int j, int) { return 0.025 + 1e-4 * ((i + j) % 5); });
        auto out = make_store<Traits, T>(d0, d1, nk, H, [](int, int, int) { return -1.; });
        st::run_single_stage(user::fused_out_f<Tag>(), backend, grid, out, in, co);
        return out;
    }

    template <class T, class Traits, class Backend>
    auto simple_hori_diff(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int k) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4. + 0.01 * k;
        };
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<Traits, T const>(d0, d1, nk, H, in_f);
        auto co = make_store<Traits, T const>(d0, d1, nk, H, [](int i, int j, int) { return 0.025 + 1e-4 * ((i + j) % 5); });
        auto out = make_store<Traits, T>(d0, d1, nk, H, [](int, int, int) { return 0.; });
        auto jb = gt::storage::builder<Traits>.template type<T const>().dimensions(d0, d1, nk).halos(H, H, 0)
                      .template selector<0, 1, 0>();
        auto cro = jb.initializer([=](int, int j, int) { return T(1. + 0.3 * std::cos(3. * j / d1)); }).build();
        auto cru = jb.initializer([=](int, int j, int) { return T(j == 0 ? 0. : 1. - 0.2 * std::sin(2. * j / d1)); }).build();
        st::run(user::simple_hori_diff_spec<T, 1>(), backend, grid, co, in, out, cro, cru);
        return out;
    }

    template <class T, class Traits, class Backend>
    auto vert_adv(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 3;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto x = [=](int i) { return 1. * i / d0; };
        auto y = [=](int j) { return 1. * j / d1; };
        auto z = [=](int k) { return 1. * k / nk; };
        fun_t u_stage_f = [=](int i, int j, int) {
            double t = x(i) + y(j);
            return 7 + std::cos(M_PI * t) + std::sin(2 * M_PI * t);
        };
        fun_t wcon_f = [=](int i, int j, int k) {
            return 2e-4 * (-1.07 + (2 + std::cos(M_PI * (x(i) + z(k))) + std::cos(M_PI * y(j))) / 2);
        };
        fun_t utens_f = [=](int i, int j, int k) {
            return 3e-6 * (-1.0235 + (2. + std::cos(M_PI * (x(i) + y(j))) + std::cos(M_PI * y(j) * z(k))) / 2);
        };
        fun_t utens_stage_f = [=](int i, int j, int k) {
            double t = x(i) + y(j);
            return 7 + 1.25 * (2. + std::cos(M_PI * t) + std::sin(2 * M_PI * t)) + .1 * k;
        };
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, user::va_axis_t(nk));
        auto utens_stage = make_store<Traits, T>(d0, d1, nk, H, utens_stage_f);
        auto u_stage = make_store<Traits, T>(d0, d1, nk, H, u_stage_f);
        auto wcon = make_store<Traits, T>(d0, d1, nk, H, wcon_f);
        auto u_pos = make_store<Traits, T>(d0, d1, nk, H, u_stage_f);
        auto utens = make_store<Traits, T>(d0, d1, nk, H, utens_f);
        st::run(user::vert_adv_spec<T, 1>(), backend, grid, utens_stage, u_stage, wcon, u_pos, utens,
            st::global_parameter(T(3. / 20.)));
        return utens_stage;
    }

    template <class Traits, class Backend>
    auto tridiagonal(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 0;
        auto hh = ij_halos(ni, nj, H);
        auto grid = st::make_grid(hh.first, hh.second, user::td_axis_t(nk));
        auto mk = [&](fun_t f) { return make_store<Traits, double>(ni, nj, nk, H, f); };
        auto out = mk([](int, int, int) { return 0.; });
        st::run(user::tridiagonal_spec<1>(), backend, grid, mk([](int, int, int) { return -1.; }),
            mk([](int, int, int) { return 3.; }), mk([](int, int, int) { return 1.; }),
            mk([=](int i, int, int k) { return (k == 0 ? 4. : k == nk - 1 ? 2. : 3.) + 1e-3 * i; }), out);
        return out;
    }

    // ------------------------------------------------------------------ k-cache patterns
    using kc_axis_t = st::axis<1, st::axis_config::offset_limit<3>>;
    using kc_full_t = kc_axis_t::full_interval;

    // out(k) = in(k-1) + in(k) + in(k+1), clipped at both ends: a filled window [-1, 1]
    struct shifted_sum_f {
        using in = in_accessor<0, extent<0, 0, 0, 0, -1, 1>>;
        using out = inout_accessor<1>;
        using param_list = make_param_list<in, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(out()) = eval(in()) + eval(in(0, 0, 1));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, -1>) {
            eval(out()) = eval(in(0, 0, -1)) + eval(in()) + eval(in(0, 0, 1));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::last_level) {
            eval(out()) = eval(in(0, 0, -1)) + eval(in());
        }
    };
    template <class Execute, class Traits, class Backend>
    auto kcache_fill(Execute execute, Traits, Backend backend, int ni, int nj, int nk) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, kc_axis_t(nk));
        auto in = make_store<Traits, double>(ni, nj, nk, 0, ramp);
        auto out = make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -7.; });
        st::run(
            [=](auto in, auto out) {
                return execute().k_cached(st::cache_io_policy::fill(), in).stage(shifted_sum_f(), in, out);
            },
            backend, grid, in, out);
        return out;
    }

    // running sums through a flushed window: forward out(k) = out(k-1) + in(k), backward out(k) = out(k+1) + in(k)
    struct prefix_sum_f {
        using in = in_accessor<0>;
        using out = inout_accessor<1, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<in, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(out()) = eval(in());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, 0>) {
            eval(out()) = eval(out(0, 0, -1)) + eval(in());
        }
    };
    struct suffix_sum_f {
        using in = in_accessor<0>;
        using out = inout_accessor<1, extent<0, 0, 0, 0, 0, 1>>;
        using param_list = make_param_list<in, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::last_level) {
            eval(out()) = eval(in());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<0, -1>) {
            eval(out()) = eval(out(0, 0, 1)) + eval(in());
        }
    };
    template <class Traits, class Backend>
    auto kcache_flush(bool forward, Traits, Backend backend, int ni, int nj, int nk) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, kc_axis_t(nk));
        auto in = make_store<Traits, double>(ni, nj, nk, 0, ramp);
        auto out = make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -7.; });
        if (forward)
            st::run(
                [](auto in, auto out) {
                    return st::execute_forward().k_cached(st::cache_io_policy::flush(), out).stage(prefix_sum_f(), in, out);
                },
                backend, grid, in, out);
        else
            st::run(
                [](auto in, auto out) {
                    return st::execute_backward().k_cached(st::cache_io_policy::flush(), out).stage(suffix_sum_f(), in, out);
                },
                backend, grid, in, out);
        return out;
    }

    // in-place running sums: the same field filled and flushed
    struct inplace_prefix_f {
        using f = inout_accessor<0, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<f>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(f()) = eval(f());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, 0>) {
            eval(f()) = eval(f()) + eval(f(0, 0, -1));
        }
    };
    struct inplace_suffix_f {
        using f = inout_accessor<0, extent<0, 0, 0, 0, 0, 1>>;
        using param_list = make_param_list<f>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::last_level) {
            eval(f()) = eval(f());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<0, -1>) {
            eval(f()) = eval(f()) + eval(f(0, 0, 1));
        }
    };
    template <class Traits, class Backend>
    auto kcache_fill_and_flush(bool forward, Traits, Backend backend, int ni, int nj, int nk) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, kc_axis_t(nk));
        auto field = make_store<Traits, double>(ni, nj, nk, 0, ramp);
        if (forward)
            st::run(
                [](auto f) {
                    return st::execute_forward()
                        .k_cached(st::cache_io_policy::fill(), st::cache_io_policy::flush(), f)
                        .stage(inplace_prefix_f(), f);
                },
                backend, grid, field);
        else
            st::run(
                [](auto f) {
                    return st::execute_backward()
                        .k_cached(st::cache_io_policy::fill(), st::cache_io_policy::flush(), f)
                        .stage(inplace_suffix_f(), f);
                },
                backend, grid, field);
        return field;
    }

    // a local window (temporary, no policy) carried through two stages of one sweep
    struct carry_f {
        using in = in_accessor<0>;
        using acc = inout_accessor<1, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<in, acc>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(acc()) = eval(in());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, 0>) {
            eval(acc()) = 0.5 * eval(acc(0, 0, -1)) + eval(in());
        }
    };
    struct emit_f {
        using acc = in_accessor<0, extent<0, 0, 0, 0, -1, 0>>;
        using out = inout_accessor<1>;
        using param_list = make_param_list<acc, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(out()) = eval(acc());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, 0>) {
            eval(out()) = eval(acc()) - eval(acc(0, 0, -1));
        }
    };
    template <class Traits, class Backend>
    auto kcache_local(Traits, Backend backend, int ni, int nj, int nk) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, kc_axis_t(nk));
        auto in = make_store<Traits, double>(ni, nj, nk, 0, ramp);
        auto out = make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -7.; });
        st::run(
            [](auto in, auto out) {
                GT_DECLARE_TMP(double, acc);
                return st::execute_forward().k_cached(acc).stage(carry_f(), in, acc).stage(emit_f(), acc, out);
            },
            backend, grid, in, out);
        return out;
    }

    // ------------------------------------------------------------------ mixed parallel multi-stage
    // two elementary intervals with different functors, an ij-cached temporary read at IJ offsets and a plain
    // temporary read at offset zero, followed by a second multi-stage that reads the plain temporary again
    using mx_axis_t = st::axis<2>;
    struct grad_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, -1, 1>>;
        using param_list = make_param_list<out, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval, mx_axis_t::get_interval<0>) {
            eval(out()) = eval(in(1, 0)) - eval(in(-1, 0));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, mx_axis_t::get_interval<1>) {
            eval(out()) = eval(in(0, 1)) - eval(in(0, -1));
        }
    };
    struct smooth_f {
        using out = inout_accessor<0>;
        using g = in_accessor<1, extent<-1, 1, -1, 1>>;
        using param_list = make_param_list<out, g>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = 0.25 * (eval(g(-1, 0)) + eval(g(1, 0)) + eval(g(0, -1)) + eval(g(0, 1)));
        }
    };
    struct combine_f {
        using out = inout_accessor<0>;
        using a = in_accessor<1>;
        using in = in_accessor<2>;
        using param_list = make_param_list<out, a, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(in()) + 2 * eval(a());
        }
    };
    struct add_f {
        using out = inout_accessor<0>;
        using a = in_accessor<1>;
        using param_list = make_param_list<out, a>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(out()) + eval(a());
        }
    };
    // CacheG = false_type: `g` is a plain temporary read at IJ offsets (CTA-private blocks in the fused path)
    template <class T, class CacheG, class Traits, class Backend>
    auto mixed_spec(CacheG, Traits, Backend backend, int ni, int nj, int nk0, int nk1) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H, nk = nk0 + nk1;
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, mx_axis_t(nk0, nk1));
        auto in = make_store<Traits, T>(
            d0, d1, nk, H, [](int i, int j, int k) { return std::sin(0.3 * i) + std::cos(0.2 * j) * (1 + 0.1 * k); });
        auto out = make_store<Traits, T>(d0, d1, nk, H, [](int, int, int) { return -3.; });
        st::run(
            [](auto in, auto out) {
                GT_DECLARE_TMP(T, g, s);
                if constexpr (!CacheG::value)
                    return st::multi_pass(st::execute_parallel()
                                              .stage(grad_f(), g, in)
                                              .stage(smooth_f(), s, g)
                                              .stage(combine_f(), out, s, in),
                        st::execute_parallel().stage(add_f(), out, s));
                else
                return st::multi_pass(st::execute_parallel()
                                          .ij_cached(g)
                                          .stage(grad_f(), g, in)
                                          .stage(smooth_f(), s, g)
                                          .stage(combine_f(), out, s, in),
                    st::execute_parallel().stage(add_f(), out, s));
            },
            backend, grid, in, out);
        return out;
    }
    // ------------------------------------------------------------------ other data stores the frontend hands to a backend
    // positional_stencil.cpp:20-43: out = i + j + k from positional<dim> "fields" (no memory behind them)
    struct position_sum_f {
        using out = inout_accessor<0>;
        using i_pos = in_accessor<1>;
        using j_pos = in_accessor<2>;
        using k_pos = in_accessor<3>;
        using param_list = make_param_list<out, i_pos, j_pos, k_pos>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(i_pos()) + 100 * eval(j_pos()) + 10000 * eval(k_pos());
        }
    };
    template <class Traits, class Backend>
    auto positional_sum(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 1;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto out = make_store<Traits, double>(d0, d1, nk, H, [](int, int, int) { return -1.; });
        st::run_single_stage(position_sum_f(), backend, grid, out, st::positional<st::dim::i>(),
            st::positional<st::dim::j>(), st::positional<st::dim::k>());
        return out;
    }

    // parallel_multistage_fusion.cpp:26-67: a temporary written by one parallel multi-stage and read one level up by the
    // next -- the two must not be fused into one launch
    using pf_axis_t = st::axis<1>;
    struct fill_level_f {
        using out = inout_accessor<0>;
        using k_pos = in_accessor<1>;
        using param_list = make_param_list<out, k_pos>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = 3 + eval(k_pos());
        }
    };
    struct copy_from_above_f {
        using in = in_accessor<0, extent<0, 0, 0, 0, 0, 1>>;
        using out = inout_accessor<1>;
        using param_list = make_param_list<in, out>;
        template <class E>
        GT_FUNCTION static void apply(E eval, pf_axis_t::full_interval::modify<0, -1>) {
            eval(out()) = eval(in(0, 0, 1));
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, pf_axis_t::full_interval::last_level) {
            eval(out()) = eval(in());
        }
    };
    template <class Traits, class Backend>
    auto parallel_multistage(Traits, Backend backend, int ni, int nj, int nk) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, pf_axis_t(nk));
        auto out = make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -1.; });
        st::run(
            [](auto out, auto k_pos) {
                GT_DECLARE_TMP(double, tmp);
                return st::multi_pass(st::execute_parallel().stage(fill_level_f(), tmp, k_pos),
                    st::execute_parallel().stage(copy_from_above_f(), tmp, out));
            },
            backend, grid, out, st::positional<st::dim::k>());
        return out;
    }

    // whole_axis_access.cpp:24-58: the k dimension of a field renamed to a fourth dimension, so that a stage can walk
    // the whole column: out(k) = sum of in over the levels below k
    struct sum_below_f {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<>, 4>;
        using k_pos = in_accessor<2>;
        using param_list = make_param_list<out, in, k_pos>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto k = eval(k_pos());
            std::decay_t<decltype(eval(out()))> res = 0;
            for (int kk = 0; kk < k; ++kk)
                res += eval(in(0, 0, 0, kk));
            eval(out()) = res;
        }
    };
    template <class Traits, class Backend>
    auto whole_axis(Traits, Backend backend, int ni, int nj, int nk) {
        using namespace gt::literals;
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<Traits, double>(ni, nj, nk, 0, ramp);
        auto out = make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -1.; });
        st::run_single_stage(sum_below_f(), backend, grid, out,
            gt::sid::rename_dimensions<st::dim::k, decltype(3_c)>(in), st::positional<st::dim::k>());
        return out;
    }

    // ------------------------------------------------------------------ several element types in one spec
    // (test_multi_types.cpp): a float tile and a double tile side by side in shared memory, an int field, a double output
    struct to_float_f {
        using t = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, 0, 0>>;
        using param_list = make_param_list<t, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(t()) = float(eval(in(1, 0)) - eval(in(-1, 0)));
        }
    };
    struct widen_f {
        using d = inout_accessor<0>;
        using t = in_accessor<1, extent<0, 0, -1, 1>>;
        using n = in_accessor<2>;
        using param_list = make_param_list<d, t, n>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(d()) = double(eval(t(0, 1))) + double(eval(t(0, -1))) + eval(n());
        }
    };
    struct gather_f {
        using out = inout_accessor<0>;
        using d = in_accessor<1, extent<-1, 1, 0, 0>>;
        using param_list = make_param_list<out, d>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(d(-1, 0)) + eval(d(1, 0));
        }
    };
    template <class Traits, class Backend>
    auto multi_types(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 3;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<Traits, double>(d0, d1, nk, H, [](int i, int j, int k) { return 0.5 * i * i + j + 0.25 * k; });
        auto n = make_store<Traits, int>(d0, d1, nk, H, [](int i, int j, int k) { return i + 2 * j + 3 * k; });
        auto out = make_store<Traits, double>(d0, d1, nk, H, [](int, int, int) { return -3.; });
        st::run(
            [](auto in, auto n, auto out) {
                GT_DECLARE_TMP(float, t);
                GT_DECLARE_TMP(double, d);
                return st::execute_parallel()
                    .ij_cached(t, d)
                    .stage(to_float_f(), t, in)
                    .stage(widen_f(), d, t, n)
                    .stage(gather_f(), out, d);
            },
            backend, grid, in, n, out);
        return out;
    }

    // ------------------------------------------------------------------ expandable_run
    // advection_pdbott_prepare_tracers.cpp:23-52: vectors of stores expanded two at a time by the frontend; the backend
    // sees one spec with two stages (and a second one with the odd tracer left over)
    struct scale_f {
        using data = inout_accessor<0>;
        using data_nnow = in_accessor<1>;
        using rho = in_accessor<2>;
        using param_list = make_param_list<data, data_nnow, rho>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(data()) = eval(rho()) * eval(data_nnow());
        }
    };
    // returns the number of tracers that differ from rho * in (bit-exact: one multiplication)
    template <class Traits, class Backend>
    int prepare_tracers(Traits, Backend backend, int ni, int nj, int nk, int tracers) {
        auto hh = ij_halos(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        using store_t = decltype(make_store<Traits, double>(ni, nj, nk, 0, fun_t()));
        std::vector<store_t> in, out;
        for (int t = 0; t < tracers; ++t) {
            out.push_back(make_store<Traits, double>(ni, nj, nk, 0, [](int, int, int) { return -1.; }));
            in.push_back(make_store<Traits, double>(ni, nj, nk, 0, [t](int i, int j, int k) { return t + 0.1 * i + 0.01 * j + k; }));
        }
        auto rho = make_store<Traits, double const>(ni, nj, nk, 0, [](int i, int, int k) { return 1.1 + 0.001 * (i + k); });
        st::expandable_run<2>(
            [](auto out, auto in, auto rho) { return st::execute_parallel().stage(scale_f(), out, in, rho); }, backend,
            grid, out, in, rho);
        int bad = 0;
        for (int t = 0; t < tracers; ++t) {
            auto o = out[t]->const_host_view();
            bool ok = true;
            for (int k = 0; k < nk; ++k)
                for (int j = 0; j < nj; ++j)
                    for (int i = 0; i < ni; ++i)
                        ok = ok && o(i, j, k) == (1.1 + 0.001 * (i + k)) * (t + 0.1 * i + 0.01 * j + k);
            bad += !ok;
        }
        return bad;
    }

    // ------------------------------------------------------------------ sweeps with IJ extents
    // forward sweep: an ij-cached difference read at j offsets feeds a local k cache; the result is flushed through
    // a second k cache into a temporary that the following parallel multi-stage reads at i offsets
    struct diff_i_f {
        using a = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, 0, 0>>;
        using param_list = make_param_list<a, in>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(a()) = eval(in(1, 0)) - eval(in(-1, 0));
        }
    };
    struct accumulate_f {
        using acc = inout_accessor<0, extent<0, 0, 0, 0, -1, 0>>;
        using c = inout_accessor<1, extent<0, 0, 0, 0, -1, 0>>;
        using a = in_accessor<2, extent<0, 0, -1, 1>>;
        using param_list = make_param_list<acc, c, a>;
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::first_level) {
            eval(acc()) = eval(a(0, 1)) + eval(a(0, -1));
            eval(c()) = eval(acc());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, kc_full_t::modify<1, 0>) {
            eval(acc()) = 0.5 * eval(acc(0, 0, -1)) + eval(a(0, 1)) + eval(a(0, -1));
            eval(c()) = eval(c(0, 0, -1)) + eval(acc());
        }
    };
    struct spread_f {
        using out = inout_accessor<0>;
        using c = in_accessor<1, extent<-1, 1, 0, 0>>;
        using param_list = make_param_list<out, c>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(out()) = eval(c(-1, 0)) + 2 * eval(c()) + eval(c(1, 0));
        }
    };
    template <class Traits, class Backend>
    auto sweep_with_extents(Traits, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 3;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto hh = ij_halos(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, kc_axis_t(nk));
        auto in = make_store<Traits, double>(
            d0, d1, nk, H, [](int i, int j, int k) { return std::sin(0.3 * i + 0.1 * k) + std::cos(0.2 * j) * (1 + 0.1 * k); });
        auto out = make_store<Traits, double>(d0, d1, nk, H, [](int, int, int) { return -3.; });
        st::run(
            [](auto in, auto out) {
                GT_DECLARE_TMP(double, a, acc, c);
                return st::multi_pass(st::execute_forward()
                                          .ij_cached(a)
                                          .k_cached(acc)
                                          .k_cached(st::cache_io_policy::flush(), c)
                                          .stage(diff_i_f(), a, in)
                                          .stage(accumulate_f(), acc, c, a),
                    st::execute_parallel().stage(spread_f(), out, c));
            },
            backend, grid, in, out);
        return out;
    }

    template <class T, class Traits, class Backend>
    auto mixed(Traits tr, Backend backend, int ni, int nj, int nk0, int nk1) {
        return mixed_spec<T>(std::true_type(), tr, backend, ni, nj, nk0, nk1);
    }
    template <class T, class Traits, class Backend>
    auto mixed_plain(Traits tr, Backend backend, int ni, int nj, int nk0, int nk1) {
        return mixed_spec<T>(std::false_type(), tr, backend, ni, nj, nk0, nk1);
    }
} // namespace cases
