/*
 * ref_driver.cpp -- builds the UNMODIFIED reference (GridTools headers under /root/reference/include) into
 * oracle/_ref/libgtref.so so that the reference's own CPU backends (stencil::cpu_ifirst, cpu_kfirst, naive)
 * can be run on this machine and on the GPU box's host cores.
 *
 * TEST INFRASTRUCTURE ONLY: used to (1) pin oracle/gt_oracle.c, (2) generate tests/golden/, (3) serve as the
 * "reference" CPU baseline in bench.py.  Never part of the product path.  No reference source is copied:
 * the headers (and the two analytic repositories of tests/regression/) are #included where they lie.
 *
 * The stencil functors below are *user code* in GridTools terms; they restate
 *   tests/regression/horizontal_diffusion.cpp:35-106, vertical_advection_dycore.cpp:32-149,
 *   tridiagonal.cpp:39-97, copy_stencil.cpp:24-36
 * with identical operand order, and are executed through the reference's frontend + backends.
 *
 * Dense exchange layout of every array crossing this C ABI: full storage box d0 x d1 x d2 (halo included),
 * i fastest:  off(i,j,k) = i + d0*(j + d1*k).  d0 = ni + 2H, d1 = nj + 2H, d2 = nk.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <omp.h>

#include <gridtools/boundaries/boundary.hpp>
#include <gridtools/boundaries/copy.hpp>
#include <gridtools/boundaries/value.hpp>
#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/cpu_ifirst.hpp>
#include <gridtools/stencil/cpu_kfirst.hpp>
#include <gridtools/stencil/global_parameter.hpp>
#include <gridtools/stencil/naive.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/cpu_kfirst.hpp>
#include <gridtools/storage/sid.hpp>

#include <horizontal_diffusion_repository.hpp>
#include <vertical_advection_repository.hpp>

namespace {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;
    using namespace gridtools::stencil;
    using namespace gridtools::stencil::cartesian;

    // ------------------------------------------------------------------ functors (user code)
    struct copy_f {
        using src = in_accessor<0>;
        using dst = inout_accessor<1>;
        using param_list = make_param_list<src, dst>;
        template <class E>
        GT_FUNCTION static void apply(E &&eval) {
            eval(dst()) = eval(src());
        }
    };

    struct hd_lap {
        using lap = inout_accessor<0>;
        using u = in_accessor<1, extent<-1, 1, -1, 1>>;
        using param_list = make_param_list<lap, u>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            using T = std::decay_t<decltype(eval(lap()))>;
            eval(lap()) = T{4} * eval(u()) - (eval(u(1, 0)) + eval(u(0, 1)) + eval(u(-1, 0)) + eval(u(0, -1)));
        }
    };
    template <int DI, int DJ>
    struct hd_flux {
        using flux = inout_accessor<0>;
        using u = in_accessor<1, extent<0, DI, 0, DJ>>;
        using lap = in_accessor<2, extent<0, DI, 0, DJ>>;
        using param_list = make_param_list<flux, u, lap>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            auto r = eval(lap(DI, DJ)) - eval(lap(0, 0));
            eval(flux()) = r * (eval(u(DI, DJ)) - eval(u(0, 0))) > 0 ? 0 : r;
        }
    };
    struct hd_out {
        using res = inout_accessor<0>;
        using u = in_accessor<1>;
        using fx = in_accessor<2, extent<-1, 0, 0, 0>>;
        using fy = in_accessor<3, extent<0, 0, -1, 0>>;
        using cf = in_accessor<4>;
        using param_list = make_param_list<res, u, fx, fy, cf>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            eval(res()) = eval(u()) - eval(cf()) * (eval(fx()) - eval(fx(-1, 0)) + eval(fy()) - eval(fy(0, -1)));
        }
    };

    // simple_hori_diff.cpp:25-61
    struct shd_wlap {
        using out = inout_accessor<0>;
        using in = in_accessor<1, extent<-1, 1, -1, 1>>;
        using crlato = in_accessor<2>;
        using crlatu = in_accessor<3>;
        using param_list = make_param_list<out, in, crlato, crlatu>;
        template <class E>
        GT_FUNCTION static void apply(E eval) {
            using T = std::decay_t<decltype(eval(out()))>;
            eval(out()) = eval(in(1, 0)) + eval(in(-1, 0)) - T{2} * eval(in()) +
                          eval(crlato()) * (eval(in(0, 1)) - eval(in())) + eval(crlatu()) * (eval(in(0, -1)) - eval(in()));
        }
    };
    struct shd_divflux {
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

    using va_axis_t = st::axis<1, st::axis_config::offset_limit<3>>;
    using va_full_t = va_axis_t::full_interval;

    struct va_forward {
        using utens_stage = in_accessor<0>;
        using wcon = in_accessor<1, extent<0, 1, 0, 0, 0, 1>>;
        using u_stage = in_accessor<2, extent<0, 0, 0, 0, -1, 1>>;
        using u_pos = in_accessor<3>;
        using utens = in_accessor<4>;
        using dtr = in_accessor<5>;
        using cc = inout_accessor<6, extent<0, 0, 0, 0, -1, 0>>;
        using dc = inout_accessor<7, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<utens_stage, wcon, u_stage, u_pos, utens, dtr, cc, dc>;

        template <class E>
        GT_FUNCTION static auto rhs(E &&eval) {
            return eval(dtr()) * eval(u_pos()) + eval(utens()) + eval(utens_stage());
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::first_level) {
            using T = std::decay_t<decltype(eval(cc()))>;
            auto gcv = T(.25) * (eval(wcon(1, 0, 1)) + eval(wcon(0, 0, 1)));
            auto cs = gcv * T(BET_M);
            auto c = gcv * T(BET_P);
            auto b = eval(dtr()) - c;
            auto corr = -cs * (eval(u_stage(0, 0, 1)) - eval(u_stage()));
            auto d = rhs(eval) + corr;
            auto inv = T(1) / b;
            eval(cc()) = c * inv;
            eval(dc()) = d * inv;
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::modify<1, -1>) {
            using T = std::decay_t<decltype(eval(cc()))>;
            auto gav = -T(.25) * (eval(wcon(1, 0, 0)) + eval(wcon(0, 0, 0)));
            auto gcv = T(.25) * (eval(wcon(1, 0, 1)) + eval(wcon(0, 0, 1)));
            auto as = gav * T(BET_M);
            auto cs = gcv * T(BET_M);
            auto a = gav * T(BET_P);
            auto c = gcv * T(BET_P);
            auto b = eval(dtr()) - a - c;
            auto corr =
                -as * (eval(u_stage(0, 0, -1)) - eval(u_stage())) - cs * (eval(u_stage(0, 0, 1)) - eval(u_stage()));
            auto d = rhs(eval) + corr;
            auto inv = T(1) / (b - eval(cc(0, 0, -1)) * a);
            eval(cc()) = c * inv;
            eval(dc()) = (d - eval(dc(0, 0, -1)) * a) * inv;
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::last_level) {
            using T = std::decay_t<decltype(eval(cc()))>;
            auto gav = -T(.25) * (eval(wcon(1, 0, 0)) + eval(wcon()));
            auto as = gav * T(BET_M);
            auto a = gav * T(BET_P);
            auto b = eval(dtr()) - a;
            auto corr = -as * (eval(u_stage(0, 0, -1)) - eval(u_stage()));
            auto d = rhs(eval) + corr;
            auto inv = T(1) / (b - eval(cc(0, 0, -1)) * a);
            eval(dc()) = (d - eval(dc(0, 0, -1)) * a) * inv;
        }
    };
    struct va_backward {
        using utens_stage = inout_accessor<0>;
        using u_pos = in_accessor<1>;
        using dtr = in_accessor<2>;
        using cc = in_accessor<3>;
        using dc = in_accessor<4>;
        using x = inout_accessor<5, extent<0, 0, 0, 0, 0, 1>>;
        using param_list = make_param_list<utens_stage, u_pos, dtr, cc, dc, x>;
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::last_level) {
            eval(utens_stage()) = eval(dtr()) * (eval(dc()) - eval(u_pos()));
            eval(x()) = eval(dc());
        }
        template <class E>
        GT_FUNCTION static void apply(E &&eval, va_full_t::modify<0, -1>) {
            auto v = eval(dc()) - eval(cc()) * eval(x(0, 0, 1));
            eval(utens_stage()) = eval(dtr()) * (v - eval(u_pos()));
            eval(x()) = v;
        }
    };

    using td_axis_t = st::axis<1>;
    using td_full_t = td_axis_t::full_interval;
    struct td_forward {
        using inf = in_accessor<0>;
        using diag = in_accessor<1>;
        using sup = inout_accessor<2, extent<0, 0, 0, 0, -1, 0>>;
        using rhs = inout_accessor<3, extent<0, 0, 0, 0, -1, 0>>;
        using param_list = make_param_list<inf, diag, sup, rhs>;
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::first_level) {
            eval(sup()) = eval(sup()) / eval(diag());
            eval(rhs()) = eval(rhs()) / eval(diag());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::modify<1, 0>) {
            eval(sup()) = eval(sup()) / (eval(diag()) - eval(sup(0, 0, -1)) * eval(inf()));
            eval(rhs()) = (eval(rhs()) - eval(inf()) * eval(rhs(0, 0, -1))) /
                          (eval(diag()) - eval(sup(0, 0, -1)) * eval(inf()));
        }
    };
    struct td_backward {
        using out = inout_accessor<0, extent<0, 0, 0, 0, 0, 1>>;
        using sup = in_accessor<1>;
        using rhs = in_accessor<2>;
        using param_list = make_param_list<out, sup, rhs>;
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::last_level) {
            eval(out()) = eval(rhs());
        }
        template <class E>
        GT_FUNCTION static void apply(E eval, td_full_t::modify<0, -1>) {
            eval(out()) = eval(rhs()) - eval(sup()) * eval(out(0, 0, 1));
        }
    };

    // ------------------------------------------------------------------ harness
    void flush_cpu_caches() {
        // same idea as tests/src/regression_main.cpp:183-192
        static std::size_t n = 1024 * 1024 * 21 / 2;
        static std::vector<double> a_(n), b_(n, 1.), c_(n, 2.);
        double *a = a_.data(), *b = b_.data(), *c = c_.data();
#pragma omp parallel for
        for (std::size_t i = 0; i < n; i++)
            a[i] = b[i] * c[i];
    }

    template <class Traits, class T>
    auto make_store(int d0, int d1, int d2, int halo, const void *dense) {
        auto b = gt::storage::builder<Traits>.template type<T>().dimensions(d0, d1, d2).halos(halo, halo, 0);
        using V = std::remove_const_t<T>;
        const V *p = static_cast<const V *>(dense);
        return b
            .initializer([=](int i, int j, int k) { return p ? p[i + (int64_t)d0 * (j + (int64_t)d1 * k)] : V(0); })
            .build();
    }
    template <class T, class Store>
    void read_store(Store const &s, int d0, int d1, int d2, void *dense) {
        T *p = static_cast<T *>(dense);
        auto v = s->const_host_view();
        for (int k = 0; k < d2; ++k)
            for (int j = 0; j < d1; ++j)
                for (int i = 0; i < d0; ++i)
                    p[i + (int64_t)d0 * (j + (int64_t)d1 * k)] = v(i, j, k);
    }

    template <class Comp>
    void timed(Comp &&comp, int nrep, int flush, double *times) {
        comp();
        for (int r = 0; r < nrep; ++r) {
            if (flush)
                flush_cpu_caches();
            double t0 = omp_get_wtime();
            comp();
            times[r] = omp_get_wtime() - t0;
        }
    }

    inline auto ij_grid(int d0, int d1, int halo) {
        auto h = [&](int d) { return gt::halo_descriptor(halo, halo, halo, d - halo - 1, d); };
        return std::make_pair(h(d0), h(d1));
    }

    template <class Backend, class Traits, class T>
    int run_copy(int ni, int nj, int nk, const void *const *in, void *const *out, int nrep, int flush, double *times) {
        auto src = make_store<Traits, T const>(ni, nj, nk, 0, in[0]);
        auto dst = make_store<Traits, T>(ni, nj, nk, 0, nullptr);
        auto hh = ij_grid(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        timed([&] { st::run_single_stage(copy_f(), Backend(), grid, src, dst); }, nrep, flush, times);
        read_store<T>(dst, ni, nj, nk, out[0]);
        return 0;
    }

    template <class Backend, class Traits, class T>
    int run_hd(int ni, int nj, int nk, const void *const *in, void *const *out, int nrep, int flush, double *times) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto u = make_store<Traits, T const>(d0, d1, nk, H, in[0]);
        auto cf = make_store<Traits, T const>(d0, d1, nk, H, in[1]);
        auto res = make_store<Traits, T>(d0, d1, nk, H, out[0]); // start from caller's content (halo untouched)
        auto hh = ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto spec = [](auto u, auto cf, auto res) {
            GT_DECLARE_TMP(T, lap, flx, fly);
            return st::execute_parallel()
                .ij_cached(lap, flx, fly)
                .stage(hd_lap(), lap, u)
                .stage(hd_flux<1, 0>(), flx, u, lap)
                .stage(hd_flux<0, 1>(), fly, u, lap)
                .stage(hd_out(), res, u, flx, fly, cf);
        };
        timed([&] { st::run(spec, Backend(), grid, u, cf, res); }, nrep, flush, times);
        read_store<T>(res, d0, d1, nk, out[0]);
        return 0;
    }

    // in: {in, coeff, crlato[d1], crlatu[d1]} (the last two are 1-d over the storage's j, selector<0,1,0> stores)
    template <class Backend, class Traits, class T>
    int run_shd(int ni, int nj, int nk, const void *const *in, void *const *out, int nrep, int flush, double *times) {
        constexpr int H = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto u = make_store<Traits, T const>(d0, d1, nk, H, in[0]);
        auto cf = make_store<Traits, T const>(d0, d1, nk, H, in[1]);
        auto res = make_store<Traits, T>(d0, d1, nk, H, out[0]);
        const T *po = static_cast<const T *>(in[2]), *pu = static_cast<const T *>(in[3]);
        auto jb = gt::storage::builder<Traits>.template type<T const>().dimensions(d0, d1, nk).halos(H, H, 0)
                      .template selector<0, 1, 0>();
        auto crlato = jb.initializer([=](int, int j, int) { return po[j]; }).build();
        auto crlatu = jb.initializer([=](int, int j, int) { return pu[j]; }).build();
        auto hh = ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, st::axis<1>(nk));
        auto spec = [](auto coeff, auto in, auto out, auto crlato, auto crlatu) {
            GT_DECLARE_TMP(T, lap);
            return st::execute_parallel()
                .ij_cached(lap)
                .stage(shd_wlap(), lap, in, crlato, crlatu)
                .stage(shd_divflux(), out, in, lap, crlato, coeff);
        };
        timed([&] { st::run(spec, Backend(), grid, cf, u, res, crlato, crlatu); }, nrep, flush, times);
        read_store<T>(res, d0, d1, nk, out[0]);
        return 0;
    }

    template <class Backend, class Traits, class T>
    int run_va(int ni, int nj, int nk, const void *const *in, void *const *out, double dtr_stage, int nrep, int flush,
        double *times) {
        constexpr int H = 3;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        // in: utens_stage_in, u_stage, wcon, u_pos, utens.  The stencil updates utens_stage in place, so for
        // nrep > 0 the timed repetitions keep transforming it (timing only); the returned field is from a
        // single application on fresh inputs.
        auto u_stage = make_store<Traits, T>(d0, d1, nk, H, in[1]);
        auto wcon = make_store<Traits, T>(d0, d1, nk, H, in[2]);
        auto u_pos = make_store<Traits, T>(d0, d1, nk, H, in[3]);
        auto utens = make_store<Traits, T>(d0, d1, nk, H, in[4]);
        auto dtr = st::global_parameter(T(dtr_stage));
        auto hh = ij_grid(d0, d1, H);
        auto grid = st::make_grid(hh.first, hh.second, va_axis_t(nk));
        auto spec = [](auto utens_stage, auto u_stage, auto wcon, auto u_pos, auto utens, auto dtr) {
            GT_DECLARE_TMP(T, ccol, dcol, data_col);
            return st::multi_pass(st::execute_forward()
                                      .k_cached(st::cache_io_policy::flush(), ccol, dcol)
                                      .k_cached(st::cache_io_policy::fill(), u_stage)
                                      .stage(va_forward(), utens_stage, wcon, u_stage, u_pos, utens, dtr, ccol, dcol),
                st::execute_backward().k_cached(data_col).stage(
                    va_backward(), utens_stage, u_pos, dtr, ccol, dcol, data_col));
        };
        {
            auto utens_stage = make_store<Traits, T>(d0, d1, nk, H, in[0]);
            st::run(spec, Backend(), grid, utens_stage, u_stage, wcon, u_pos, utens, dtr);
            read_store<T>(utens_stage, d0, d1, nk, out[0]);
        }
        if (nrep > 0) {
            auto utens_stage = make_store<Traits, T>(d0, d1, nk, H, in[0]);
            timed([&] { st::run(spec, Backend(), grid, utens_stage, u_stage, wcon, u_pos, utens, dtr); },
                nrep,
                flush,
                times);
        }
        return 0;
    }

    template <class Backend, class Traits, class T>
    int run_td(int ni, int nj, int nk, const void *const *in, void *const *out, int nrep, int flush, double *times) {
        // in: inf, diag, sup, rhs ; out: out, sup', rhs'
        auto inf = make_store<Traits, T>(ni, nj, nk, 0, in[0]);
        auto diag = make_store<Traits, T>(ni, nj, nk, 0, in[1]);
        auto sup = make_store<Traits, T>(ni, nj, nk, 0, in[2]);
        auto rhs = make_store<Traits, T>(ni, nj, nk, 0, in[3]);
        auto x = make_store<Traits, T>(ni, nj, nk, 0, nullptr);
        auto hh = ij_grid(ni, nj, 0);
        auto grid = st::make_grid(hh.first, hh.second, td_axis_t(nk));
        st::run(
            [](auto inf, auto diag, auto sup, auto rhs, auto x) {
                return st::multi_pass(st::execute_forward().stage(td_forward(), inf, diag, sup, rhs),
                    st::execute_backward().stage(td_backward(), x, sup, rhs));
            },
            Backend(),
            grid,
            inf,
            diag,
            sup,
            rhs,
            x);
        read_store<T>(x, ni, nj, nk, out[0]);
        if (out[1])
            read_store<T>(sup, ni, nj, nk, out[1]);
        if (out[2])
            read_store<T>(rhs, ni, nj, nk, out[2]);
        (void)nrep, (void)flush, (void)times;
        return 0;
    }

    using be_ifirst = st::cpu_ifirst<>;
    using be_kfirst = st::cpu_kfirst<>;
    using be_naive = st::naive;
    using tr_ifirst = gt::storage::cpu_ifirst;
    using tr_kfirst = gt::storage::cpu_kfirst;
} // namespace

namespace {
    struct mask_predicate {
        const int *mask;
        template <class D>
        bool operator()(D) const {
            return mask[((int)D::i + 1) + 3 * ((int)D::j + 1) + 9 * ((int)D::k + 1)] != 0;
        }
    };
} // namespace

extern "C" {

enum { GTREF_COPY = 0, GTREF_HORI_DIFF = 1, GTREF_VERT_ADV = 2, GTREF_TRIDIAGONAL = 3, GTREF_SIMPLE_HORI_DIFF = 4 };
enum { GTREF_CPU_IFIRST = 0, GTREF_CPU_KFIRST = 1, GTREF_NAIVE = 2 };

int gtref_num_threads() { return omp_get_max_threads(); }

/* Runs `stencil` on `backend` (ni,nj,nk = compute domain).  Returns 0 on success, 1 for an unsupported combination.
 * times[nrep] receives seconds of each timed repetition (after one warm-up); flush != 0 streams 3x88 MB through the
 * caches before each repetition. */
int gtref_run(int stencil, int backend, int elem_size, int ni, int nj, int nk, const void *const *in, void *const *out,
    double scalar, int nrep, int flush, double *times) {
#define DISPATCH(FN, T, ...)                                                   \
    switch (backend) {                                                         \
    case GTREF_CPU_IFIRST:                                                     \
        return FN<be_ifirst, tr_ifirst, T>(__VA_ARGS__);                       \
    case GTREF_CPU_KFIRST:                                                     \
        return FN<be_kfirst, tr_kfirst, T>(__VA_ARGS__);                       \
    case GTREF_NAIVE:                                                          \
        return FN<be_naive, tr_ifirst, T>(__VA_ARGS__);                        \
    default:                                                                   \
        return 1;                                                              \
    }
    switch (stencil) {
    case GTREF_COPY:
        if (elem_size == 8) {
            DISPATCH(run_copy, double, ni, nj, nk, in, out, nrep, flush, times)
        }
        return 1;
    case GTREF_HORI_DIFF:
        if (elem_size == 8) {
            DISPATCH(run_hd, double, ni, nj, nk, in, out, nrep, flush, times)
        } else if (elem_size == 4) {
            DISPATCH(run_hd, float, ni, nj, nk, in, out, nrep, flush, times)
        }
        return 1;
    case GTREF_VERT_ADV:
        if (elem_size == 8) {
            DISPATCH(run_va, double, ni, nj, nk, in, out, scalar, nrep, flush, times)
        } else if (elem_size == 4) {
            DISPATCH(run_va, float, ni, nj, nk, in, out, scalar, nrep, flush, times)
        }
        return 1;
    case GTREF_TRIDIAGONAL:
        if (elem_size == 8) {
            DISPATCH(run_td, double, ni, nj, nk, in, out, nrep, flush, times)
        }
        return 1;
    case GTREF_SIMPLE_HORI_DIFF:
        if (elem_size == 8) {
            DISPATCH(run_shd, double, ni, nj, nk, in, out, nrep, flush, times)
        } else if (elem_size == 4) {
            DISPATCH(run_shd, float, ni, nj, nk, in, out, nrep, flush, times)
        }
        return 1;
    }
    return 1;
#undef DISPATCH
}

/* Analytic repositories of the reference's regression tests, evaluated on the full storage box d0 x d1 x d2
 * (halo included; d0 = ni + 4 etc. exactly like test_environment<2>::d()).  `out` is only meaningful on the
 * interior.  horizontal_diffusion_repository.hpp:32-66. */
void gtref_repo_hori_diff(int d0, int d1, int d2, double *in, double *coeff, double *out) {
    gt::horizontal_diffusion_repository repo(d0, d1, d2);
    for (int k = 0; k < d2; ++k)
        for (int j = 0; j < d1; ++j)
            for (int i = 0; i < d0; ++i) {
                int64_t o = i + (int64_t)d0 * (j + (int64_t)d1 * k);
                in[o] = repo.in(i, j, k);
                coeff[o] = repo.coeff(i, j, k);
                bool interior = i >= 2 && i < d0 - 2 && j >= 2 && j < d1 - 2;
                out[o] = interior ? repo.out(i, j, k) : 0.;
            }
}

/* horizontal_diffusion_repository.hpp:32-46,68-79: in, coeff, crlato(j), crlatu(j) and out_simple (interior). */
void gtref_repo_simple_hori_diff(int d0, int d1, int d2, double *in, double *coeff, double *crlato, double *crlatu,
    double *out_simple) {
    gt::horizontal_diffusion_repository repo(d0, d1, d2);
    for (int j = 0; j < d1; ++j) {
        crlato[j] = repo.crlato(0, j, 0);
        crlatu[j] = repo.crlatu(0, j, 0);
    }
    for (int k = 0; k < d2; ++k)
        for (int j = 0; j < d1; ++j)
            for (int i = 0; i < d0; ++i) {
                int64_t o = i + (int64_t)d0 * (j + (int64_t)d1 * k);
                in[o] = repo.in(i, j, k);
                coeff[o] = repo.coeff(i, j, k);
                bool interior = i >= 2 && i < d0 - 2 && j >= 2 && j < d1 - 2;
                out_simple[o] = interior ? repo.out_simple(i, j, k) : 0.;
            }
}

/* The reference's boundary<BoundaryFunction, gcl::cpu, Predicate>::apply (boundaries/boundary.hpp:57-72) on cpu_ifirst
 * stores built from dense boxes [d2][d1][d0]; kind 0 = value_boundary<double>(value), 1 = copy_boundary (the last field
 * is the source); halo[15] = (minus, plus, begin, end, total) x 3; mask[27] = the predicate.  1 to 3 fields. */
int gtref_boundary(int kind, double value, const int halo[15], const int mask[27], int d0, int d1, int d2,
    double *const *fields, int n_fields) {
    namespace bd = gt::boundaries;
    if (n_fields < 1 || n_fields > 3)
        return 1;
    gt::array<gt::halo_descriptor, 3> hd{gt::halo_descriptor(halo[0], halo[1], halo[2], halo[3], halo[4]),
        gt::halo_descriptor(halo[5], halo[6], halo[7], halo[8], halo[9]),
        gt::halo_descriptor(halo[10], halo[11], halo[12], halo[13], halo[14])};
    auto mk = [&](const double *p) {
        return gt::storage::builder<gt::storage::cpu_ifirst>.type<double>().dimensions(d0, d1, d2)
            .initializer([=](int i, int j, int k) { return p[i + (int64_t)d0 * (j + (int64_t)d1 * k)]; })
            .build();
    };
    auto a = mk(fields[0]);
    auto b = mk(fields[n_fields > 1 ? 1 : 0]);
    auto c = mk(fields[n_fields > 2 ? 2 : 0]);
    mask_predicate pred{mask};
    if (kind == 0) {
        auto bc = bd::make_boundary<gt::gcl::cpu>(hd, bd::value_boundary<double>(value), pred);
        if (n_fields == 1)
            bc.apply(a);
        else if (n_fields == 2)
            bc.apply(a, b);
        else
            bc.apply(a, b, c);
    } else if (kind == 1) {
        auto bc = bd::make_boundary<gt::gcl::cpu>(hd, bd::copy_boundary(), pred);
        if (n_fields == 2)
            bc.apply(a, b);
        else if (n_fields == 3)
            bc.apply(a, b, c);
        else
            return 1;
    } else
        return 1;
    read_store<double>(a, d0, d1, d2, fields[0]);
    if (n_fields > 1)
        read_store<double>(b, d0, d1, d2, fields[1]);
    if (n_fields > 2)
        read_store<double>(c, d0, d1, d2, fields[2]);
    return 0;
}

/* vertical_advection_repository.hpp:70-151; fields[5] = utens_stage_in, u_stage, wcon, u_pos, utens. */
void gtref_repo_vert_adv(int d0, int d1, int d2, double *const *fields, double *utens_stage_out, double *dtr_stage) {
    gt::vertical_advection_repository repo(d0, d1, d2);
    *dtr_stage = repo.dtr_stage;
    for (int k = 0; k < d2; ++k)
        for (int j = 0; j < d1; ++j)
            for (int i = 0; i < d0; ++i) {
                int64_t o = i + (int64_t)d0 * (j + (int64_t)d1 * k);
                fields[0][o] = repo.utens_stage_in(i, j, k);
                fields[1][o] = repo.u_stage(i, j, k);
                fields[2][o] = repo.wcon(i, j, k);
                fields[3][o] = repo.u_pos(i, j, k);
                fields[4][o] = repo.utens(i, j, k);
            }
    if (utens_stage_out)
        for (int j = 3; j < d1 - 3; ++j)
            for (int i = 3; i < d0 - 3; ++i)
                for (int k = 0; k < d2; ++k)
                    utens_stage_out[i + (int64_t)d0 * (j + (int64_t)d1 * k)] = repo.utens_stage_out(i, j, k);
}
}
