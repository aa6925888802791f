// b200_regression.cu -- the stencil::b200 backend tag driven through GridTools' own frontend (stencil::run /
// run_single_stage, storage::builder, make_grid), compared in the same process with the reference's cpu_ifirst
// backend on the same inputs.  Tag 0 functors are registered to the named sm_100a kernels, tag 1 functors are not and
// exercise the generic paths (more of those in b200_generic.cu).  Built here against /root/reference/include; the binary travels to the GPU box.
//
//   b200_regression            -> prints one line per case and "ALL PASSED" / "FAILED"; exit code 0 / 1
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/cpu_ifirst.hpp>
#include <gridtools/stencil/global_parameter.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/gpu.hpp>
#include <gridtools/storage/sid.hpp>

#include <gridtools/boundaries/boundary.hpp>
#include <gridtools/boundaries/copy.hpp>
#include <gridtools/boundaries/value.hpp>

#include <gtb200/boundaries/boundary.hpp>
#include <gtb200/stencil/b200.hpp>

#include "functors.hpp"

GTB200_REGISTER_SPEC(gtb200::kernel::copy, user::copy_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, user::lap_f<0>, user::flx_f<0>, user::fly_f<0>, user::out_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::simple_hori_diff, user::wlap_f<0>, user::divflux_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::vert_adv, user::va_forward_f<0>, user::va_backward_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::tridiagonal, user::td_forward_f<0>, user::td_backward_f<0>);

namespace {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;
    using fun_t = std::function<double(int, int, int)>;

    int g_failed = 0;

    template <class Traits, class T>
    auto make_store(int d0, int d1, int d2, int halo, fun_t f) {
        return gt::storage::builder<Traits>.template type<T>().dimensions(d0, d1, d2).halos(halo, halo, 0)
            .initializer([f](int i, int j, int k) { return T(f(i, j, k)); })
            .build();
    }

    auto make_ij_grid(int d0, int d1, int halo) {
        auto h = [&](int d) { return gt::halo_descriptor(halo, halo, halo, d - halo - 1, d); };
        return std::make_pair(h(d0), h(d1));
    }

    // tests/include/verifier.hpp:26-52 on the compute domain
    template <class A, class B>
    bool verify(const char *name, A const &a, B const &b, int d0, int d1, int d2, int halo, double tol) {
        auto va = a->const_host_view();
        auto vb = b->const_host_view();
        double worst = 0;
        long bad = 0;
        for (int k = 0; k < d2; ++k)
            for (int j = halo; j < d1 - halo; ++j)
                for (int i = halo; i < d0 - halo; ++i) {
                    double x = va(i, j, k), y = vb(i, j, k);
                    double d = std::fabs(x - y), s = std::fmax(std::fabs(x), std::fabs(y));
                    double rel = s > 0 ? d / s : 0;
                    if (tol == 0 ? x != y : !(d < tol || rel < tol))
                        ++bad;
                    if (rel > worst && d >= tol)
                        worst = rel;
                }
        std::printf("%-44s %s (mismatches %ld, worst rel %.3g)\n", name, bad ? "FAILED" : "ok", bad, worst);
        if (bad)
            ++g_failed;
        return bad == 0;
    }

    template <class T, int Tag>
    void test_copy(int ni, int nj, int nk) {
        fun_t f = [](int i, int j, int k) { return i + 100. * j + 1e4 * k + 0.5; };
        auto grid_h = make_ij_grid(ni, nj, 0);
        auto grid = st::make_grid(grid_h.first, grid_h.second, st::axis<1>(nk));
        auto in = make_store<gt::storage::gpu, T const>(ni, nj, nk, 0, f);
        auto out = make_store<gt::storage::gpu, T>(ni, nj, nk, 0, [](int, int, int) { return -1.; });
        st::run_single_stage(user::copy_f<Tag>(), st::b200<>(), grid, in, out);
        auto ref = make_store<gt::storage::cpu_ifirst, T>(ni, nj, nk, 0, f);
        std::string name = std::string("copy ") + (Tag ? "generic" : "named") + (sizeof(T) == 8 ? " f64" : " f32");
        verify(name.c_str(), out, ref, ni, nj, nk, 0, 0.);
    }

    template <class T, int Tag>
    void test_hori_diff(int ni, int nj, int nk) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int k) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4. + 0.01 * k;
        };
        fun_t co_f = [](int i, int j, int) { return 0.025 + 1e-4 * ((i + j) % 5); };
        auto hh = make_ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto in = make_store<gt::storage::gpu, T const>(d0, d1, nk, H, in_f);
        auto co = make_store<gt::storage::gpu, T const>(d0, d1, nk, H, co_f);
        auto out = make_store<gt::storage::gpu, T>(d0, d1, nk, H, [](int, int, int) { return 0.; });
        st::run(user::hori_diff_spec<T, Tag>(), st::b200<>(), grid, in, co, out);
        auto in_r = make_store<gt::storage::cpu_ifirst, T const>(d0, d1, nk, H, in_f);
        auto co_r = make_store<gt::storage::cpu_ifirst, T const>(d0, d1, nk, H, co_f);
        auto out_r = make_store<gt::storage::cpu_ifirst, T>(d0, d1, nk, H, [](int, int, int) { return 0.; });
        st::run(user::hori_diff_spec<T, 1>(), st::cpu_ifirst<>(), grid, in_r, co_r, out_r);
        std::string name = std::string("horizontal_diffusion ") + (Tag ? "generic" : "named") +
                           (sizeof(T) == 8 ? " f64 " : " f32 ") + std::to_string(ni) + "x" + std::to_string(nj) + "x" +
                           std::to_string(nk);
        verify(name.c_str(), out, out_r, d0, d1, nk, H, sizeof(T) == 8 ? 1e-12 : 1e-5);
    }

    template <class T, int Tag>
    void test_simple_hori_diff(int ni, int nj, int nk) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int k) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4. + 0.01 * k;
        };
        fun_t co_f = [](int i, int j, int) { return 0.025 + 1e-4 * ((i + j) % 5); };
        fun_t cro_f = [=](int, int j, int) { return 1. + 0.3 * std::cos(3. * j / d1); };
        fun_t cru_f = [=](int, int j, int) { return j == 0 ? 0. : 1. - 0.2 * std::sin(2. * j / d1); };
        auto hh = make_ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto run_on = [&](auto traits, auto backend, auto tag) {
            using traits_t = decltype(traits);
            auto in = make_store<traits_t, T const>(d0, d1, nk, H, in_f);
            auto co = make_store<traits_t, T const>(d0, d1, nk, H, co_f);
            auto out = make_store<traits_t, T>(d0, d1, nk, H, [](int, int, int) { return 0.; });
            auto jb = gt::storage::builder<traits_t>.template type<T const>().dimensions(d0, d1, nk).halos(H, H, 0)
                          .template selector<0, 1, 0>();
            auto cro = jb.initializer([=](int i, int j, int k) { return T(cro_f(i, j, k)); }).build();
            auto cru = jb.initializer([=](int i, int j, int k) { return T(cru_f(i, j, k)); }).build();
            st::run(user::simple_hori_diff_spec<T, decltype(tag)::value>(), backend, grid, co, in, out, cro, cru);
            return out;
        };
        auto got = run_on(gt::storage::gpu(), st::b200<>(), std::integral_constant<int, Tag>());
        auto ref = run_on(gt::storage::cpu_ifirst(), st::cpu_ifirst<>(), std::integral_constant<int, 1>());
        std::string name = std::string("simple_hori_diff ") + (Tag ? "generic" : "named") +
                           (sizeof(T) == 8 ? " f64 " : " f32 ") + std::to_string(ni) + "x" + std::to_string(nj) + "x" +
                           std::to_string(nk);
        verify(name.c_str(), got, ref, d0, d1, nk, H, sizeof(T) == 8 ? 1e-12 : 1e-5);
    }

    template <class T, int Tag>
    void test_vert_adv(int ni, int nj, int nk) {
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
        const T dtr = T(3. / 20.);
        auto hh = make_ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, user::va_axis_t(nk));
        auto run_on = [&](auto traits, auto backend, auto tag) {
            using traits_t = decltype(traits);
            auto utens_stage = make_store<traits_t, T>(d0, d1, nk, H, utens_stage_f);
            auto u_stage = make_store<traits_t, T>(d0, d1, nk, H, u_stage_f);
            auto wcon = make_store<traits_t, T>(d0, d1, nk, H, wcon_f);
            auto u_pos = make_store<traits_t, T>(d0, d1, nk, H, u_stage_f);
            auto utens = make_store<traits_t, T>(d0, d1, nk, H, utens_f);
            st::run(user::vert_adv_spec<T, decltype(tag)::value>(), backend, grid, utens_stage, u_stage, wcon, u_pos,
                utens, st::global_parameter(dtr));
            return utens_stage;
        };
        auto got = run_on(gt::storage::gpu(), st::b200<>(), std::integral_constant<int, Tag>());
        auto ref = run_on(gt::storage::cpu_ifirst(), st::cpu_ifirst<>(), std::integral_constant<int, 1>());
        std::string name = std::string("vertical_advection_dycore ") + (Tag ? "generic" : "named") +
                           (sizeof(T) == 8 ? " f64 " : " f32 ") + std::to_string(ni) + "x" + std::to_string(nj) + "x" +
                           std::to_string(nk);
        verify(name.c_str(), got, ref, d0, d1, nk, H, sizeof(T) == 8 ? 1e-12 : 1e-4);
    }

    // boundaries/boundary.hpp:57-72: the reference's boundary<..., gcl::cpu>::apply on cpu_ifirst stores against
    // gtb200::boundaries::boundary on storage::gpu stores (raw target pointers), whole storages compared bit for bit.
    struct all_predicate { // both call forms: the reference passes a direction type, gtb200 three ints
        template <class D>
        bool operator()(D) const {
            return true;
        }
        bool operator()(int, int, int) const { return true; }
    };
    struct i_minus_predicate { // applies the condition only on the directions that point towards -i
        template <class D>
        bool operator()(D) const {
            return D::i == gt::boundaries::minus_;
        }
        bool operator()(int ei, int, int) const { return ei < 0; }
    };

    template <class Pred>
    void test_boundary(const char *what, Pred pred) {
        const int d0 = 21, d1 = 13, d2 = 6;
        namespace rb = gt::boundaries;
        namespace bd = gtb200::boundaries;
        fun_t fa = [](int i, int j, int k) { return 1. + i + 100. * j + 1e4 * k; };
        fun_t fb = [](int i, int j, int k) { return -2. - i - 50. * j - 1e3 * k; };
        gt::array<gt::halo_descriptor, 3> hd{gt::halo_descriptor(2, 3, 2, d0 - 4, d0),
            gt::halo_descriptor(1, 2, 1, d1 - 3, d1), gt::halo_descriptor(1, 1, 1, d2 - 2, d2)};
        auto mk = [&](auto traits, fun_t f) {
            using traits_t = decltype(traits);
            return gt::storage::builder<traits_t>.template type<double>().dimensions(d0, d1, d2)
                .initializer([f](int i, int j, int k) { return f(i, j, k); }).build();
        };
        for (int kind = 0; kind < 2; ++kind) {
            auto ra = mk(gt::storage::cpu_ifirst(), fa), rb_ = mk(gt::storage::cpu_ifirst(), fb);
            auto ga = mk(gt::storage::gpu(), fa), gb = mk(gt::storage::gpu(), fb);
            // storage::gpu pads the i-length: the device layout's total length in i is the j stride
            const int p0 = (int)ga->strides()[1];
            std::array<gtb_halo_desc, 3> h = {{{2, 3, 2, d0 - 4, p0}, {1, 2, 1, d1 - 3, d1}, {1, 1, 1, d2 - 2, d2}}};
            if (kind == 0) {
                rb::make_boundary<gt::gcl::cpu>(hd, rb::value_boundary<double>(7.25), pred).apply(ra, rb_);
                bd::make_boundary(h, bd::value_boundary<double>(7.25), pred).apply(ga->get_target_ptr(), gb->get_target_ptr());
            } else {
                rb::make_boundary<gt::gcl::cpu>(hd, rb::copy_boundary(), pred).apply(ra, rb_);
                bd::make_boundary(h, bd::copy_boundary(), pred).apply(ga->get_target_ptr(), gb->get_target_ptr());
            }
            cudaDeviceSynchronize();
            std::string name = std::string("boundary ") + (kind ? "copy " : "value ") + what;
            verify((name + " field 0").c_str(), ga, ra, d0, d1, d2, 0, 0);
            verify((name + " field 1").c_str(), gb, rb_, d0, d1, d2, 0, 0);
        }
    }

    template <int Tag>
    void test_tridiagonal(int ni, int nj, int nk) {
        auto hh = make_ij_grid(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, user::td_axis_t(nk));
        fun_t rhs_f = [=](int, int, int k) { return k == 0 ? 4. : k == nk - 1 ? 2. : 3.; };
        auto mk = [&](fun_t f) { return make_store<gt::storage::gpu, double>(ni, nj, nk, 0, f); };
        auto out = mk([](int, int, int) { return 0.; });
        st::run(user::tridiagonal_spec<Tag>(), st::b200<>(), grid, mk([](int, int, int) { return -1.; }),
            mk([](int, int, int) { return 3.; }), mk([](int, int, int) { return 1.; }), mk(rhs_f), out);
        auto ones = make_store<gt::storage::cpu_ifirst, double>(ni, nj, nk, 0, [](int, int, int) { return 1.; });
        std::string name = std::string("tridiagonal ") + (Tag ? "generic" : "named") + " (solution == 1)";
        verify(name.c_str(), out, ones, ni, nj, nk, 0, 1e-14); // tridiagonal.cpp:97
    }

} // namespace

int main() {
    try {
        test_copy<double, 0>(40, 13, 7);
        test_copy<float, 0>(64, 8, 3);
        test_copy<double, 1>(23, 11, 5);
        test_hori_diff<double, 0>(12, 33, 61); // test_environment sizes of the reference suite
        test_hori_diff<double, 0>(23, 11, 43);
        test_hori_diff<double, 0>(128, 128, 80); // BASELINE.json configs[0]
        test_hori_diff<float, 0>(70, 19, 5);
        test_hori_diff<double, 1>(23, 11, 7); // generic path
        test_hori_diff<float, 1>(33, 9, 3);
        test_simple_hori_diff<double, 0>(12, 33, 61);
        test_simple_hori_diff<double, 0>(70, 19, 5);
        test_simple_hori_diff<float, 0>(23, 11, 7);
        test_simple_hori_diff<double, 1>(23, 11, 7); // generic path with j-only fields
        test_vert_adv<double, 0>(12, 33, 61);
        test_vert_adv<double, 0>(23, 11, 43);
        test_vert_adv<double, 0>(256, 256, 80); // BASELINE.json configs[1]
        test_vert_adv<float, 0>(40, 9, 20);
        test_vert_adv<double, 1>(23, 11, 43); // generic path: forward/backward sweeps, k-cached temporaries
        test_boundary("all directions", all_predicate());
        test_boundary("-i directions", i_minus_predicate());
        test_tridiagonal<0>(12, 33, 6);
        test_tridiagonal<0>(23, 11, 6);
        test_tridiagonal<1>(23, 11, 6);
    } catch (std::exception const &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::puts(g_failed ? "FAILED" : "ALL PASSED");
    return g_failed ? 1 : 0;
}
