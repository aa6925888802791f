// gpu_incumbent.cu -- the reference's OWN CUDA backend (stencil::gpu<>, unmodified headers, recompiled for sm_100a) and
// the stencil::b200<> tag, both driven through GridTools' frontend on the same device storages and timed with CUDA
// events.  This is "the number to beat" of SURVEY.md section 8d: horizontal_diffusion and vertical_advection_dycore
// at 256x256x80 fp64 (perftests sizes of tests/regression/horizontal_diffusion.cpp:125,
// vertical_advection_dycore.cpp:128), three rotating field sets so that no run finds its inputs in L2.
//
//   gpu_incumbent [ni nj nk]   -> one line per (stencil, backend): microseconds per run, Mpts/s, GB/s algorithmic
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <tuple>

#include <cuda_runtime.h>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/global_parameter.hpp>
#include <gridtools/stencil/gpu.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/gpu.hpp>
#include <gridtools/storage/sid.hpp>

#include <gtb200/stencil/b200.hpp>

#include "functors.hpp"

GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, user::lap_f<0>, user::flx_f<0>, user::fly_f<0>, user::out_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::vert_adv, user::va_forward_f<0>, user::va_backward_f<0>);

namespace {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;
    using fun_t = std::function<double(int, int, int)>;

    template <class Traits, class T>
    auto make_store_on(int d0, int d1, int d2, int halo, fun_t f) {
        return gt::storage::builder<Traits>.template type<T>().dimensions(d0, d1, d2).halos(halo, halo, 0)
            .initializer([f](int i, int j, int k) { return T(f(i, j, k)); })
            .build();
    }
    template <class T>
    auto make_store(int d0, int d1, int d2, int halo, fun_t f) {
        return make_store_on<gt::storage::gpu, T>(d0, d1, d2, halo, f);
    }

    template <class F>
    double time_us(F &&run, int sets) {
        for (int s = 0; s < 12; ++s)
            run(s % sets);
        cudaDeviceSynchronize();
        cudaEvent_t a, b;
        cudaEventCreate(&a), cudaEventCreate(&b);
        const int n = 200;
        cudaEventRecord(a);
        for (int s = 0; s < n; ++s)
            run(s % sets);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms * 1e3 / n;
    }

    void report(const char *stencil, const char *backend, double us, double pts, double bytes_per_pt) {
        std::printf("%-28s %-24s %9.2f us/run %10.0f Mpts/s %8.0f GB/s algorithmic\n", stencil, backend, us, pts / us,
            pts * bytes_per_pt / us / 1e3);
    }

    template <int Tag, class Backend>
    void hori_diff(const char *name, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 2, SETS = 3;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        fun_t in_f = [=](int i, int j, int) {
            double x = 1. * i / d0, y = 1. * j / d1;
            return 5. + 8 * (2. + std::cos(M_PI * (x + 1.5 * y)) + std::sin(2 * M_PI * (x + 1.5 * y))) / 4.;
        };
        auto h = [&](int d) { return gt::halo_descriptor(H, H, H, d - H - 1, d); };
        auto grid = st::make_grid(h(d0), h(d1), st::axis<1>(nk));
        auto make_set = [&] {
            return std::make_tuple(make_store<double const>(d0, d1, nk, H, in_f),
                make_store<double const>(d0, d1, nk, H, [](int, int, int) { return 0.025; }),
                make_store<double>(d0, d1, nk, H, [](int, int, int) { return 0.; }));
        };
        auto s0 = make_set(), s1 = make_set(), s2 = make_set();
        auto call = [&](auto &f) {
            st::run(user::hori_diff_spec<double, Tag>(), backend, grid, std::get<0>(f), std::get<1>(f), std::get<2>(f));
        };
        double us = time_us([&](int s) { s == 0 ? call(s0) : (s == 1 ? call(s1) : call(s2)); }, SETS);
        report("horizontal_diffusion", name, us, 1. * ni * nj * nk, 24);
    }

    template <int Tag, class Backend>
    void vert_adv(const char *name, Backend backend, int ni, int nj, int nk) {
        constexpr int H = 3, SETS = 2;
        const int d0 = ni + 2 * H, d1 = nj + 2 * H;
        auto x = [=](int i) { return 1. * i / d0; };
        auto y = [=](int j) { return 1. * j / d1; };
        auto z = [=](int k) { return 1. * k / nk; };
        fun_t u_f = [=](int i, int j, int) {
            double t = x(i) + y(j);
            return 7 + std::cos(M_PI * t) + std::sin(2 * M_PI * t);
        };
        fun_t wcon_f = [=](int i, int j, int k) {
            return 2e-4 * (-1.07 + (2 + std::cos(M_PI * (x(i) + z(k))) + std::cos(M_PI * y(j))) / 2);
        };
        fun_t utens_f = [=](int i, int j, int k) {
            return 3e-6 * (-1.0235 + (2. + std::cos(M_PI * (x(i) + y(j))) + std::cos(M_PI * y(j) * z(k))) / 2);
        };
        fun_t us_f = [=](int i, int j, int k) {
            double t = x(i) + y(j);
            return 7 + 1.25 * (2. + std::cos(M_PI * t) + std::sin(2 * M_PI * t)) + .1 * k;
        };
        auto h = [&](int d) { return gt::halo_descriptor(H, H, H, d - H - 1, d); };
        auto grid = st::make_grid(h(d0), h(d1), user::va_axis_t(nk));
        auto make_set = [&] {
            return std::make_tuple(make_store<double>(d0, d1, nk, H, us_f), make_store<double>(d0, d1, nk, H, u_f),
                make_store<double>(d0, d1, nk, H, wcon_f), make_store<double>(d0, d1, nk, H, u_f),
                make_store<double>(d0, d1, nk, H, utens_f));
        };
        auto s0 = make_set(), s1 = make_set();
        const double dtr = 3. / 20.;
        auto call = [&](auto &f) {
            st::run(user::vert_adv_spec<double, Tag>(), backend, grid, std::get<0>(f), std::get<1>(f), std::get<2>(f),
                std::get<3>(f), std::get<4>(f), st::global_parameter(dtr));
        };
        double t = time_us([&](int s) { s == 0 ? call(s0) : call(s1); }, SETS);
        report("vertical_advection_dycore", name, t, 1. * ni * nj * nk, 48);
    }
    template <int BI, int BJ, int KB, int U = 1, bool Chain = true, int P = 0, bool L1 = false, int PP = 0, bool Stage = true,
        bool RegisterTiles = false, bool L2Hints = false>
    using fused_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
        gtb200::block_geometry<BI, BJ, KB, U, Chain, P, L1, PP, Stage, RegisterTiles, L2Hints>>;
    template <int BI, int BJ, int KB, bool Stage = true>
    using rtile_t = fused_t<BI, BJ, KB, 3, true, 4, true, 0, Stage, true>;
} // namespace

int main(int argc, char **argv) {
    const int ni = argc > 3 ? std::atoi(argv[1]) : 256, nj = argc > 3 ? std::atoi(argv[2]) : 256,
              nk = argc > 3 ? std::atoi(argv[3]) : 80;
    try {
        std::printf("# %dx%dx%d fp64, CUDA events around 200 runs, rotating field sets\n", ni, nj, nk);
        // tag 0 functors are bound to the named kernels, tag 1 functors are not: generic paths of the same tag
        using staged_t = st::b200<gtb200::default_stream, gtb200::stage_by_stage>;
#ifndef FUSED_SWEEP
        hori_diff<1>("stencil::gpu<>", st::gpu<>(), ni, nj, nk);
        hori_diff<0>("stencil::b200<> named", st::b200<>(), ni, nj, nk);
        hori_diff<1>("stencil::b200<> fused", st::b200<>(), ni, nj, nk);
        hori_diff<1>("stencil::b200<> staged", staged_t(), ni, nj, nk);
        vert_adv<1>("stencil::gpu<>", st::gpu<>(), ni, nj, nk);
        vert_adv<0>("stencil::b200<> named", st::b200<>(), ni, nj, nk);
        vert_adv<1>("stencil::b200<> fused", st::b200<>(), ni, nj, nk);
        vert_adv<1>("stencil::b200<> staged", staged_t(), ni, nj, nk);
#else
        // block geometries / sweep unroll factors of the fused generic path (make -C tests/cpp fused_timing)
        // register tiles: one thread per column and level, temporaries in registers sliding along j
        hori_diff<1>("register tiles 32x8x4", rtile_t<32, 8, 4>(), ni, nj, nk);
        hori_diff<1>("register tiles 32x16x4", rtile_t<32, 16, 4>(), ni, nj, nk);
        hori_diff<1>("register tiles 64x16x4", rtile_t<64, 16, 4>(), ni, nj, nk);
        hori_diff<1>("register tiles 64x8x4", rtile_t<64, 8, 4>(), ni, nj, nk);
        hori_diff<1>("register tiles 32x16x4 not staged", rtile_t<32, 16, 4, false>(), ni, nj, nk);
        hori_diff<1>("register tiles 128x16x2 not staged", rtile_t<128, 16, 2, false>(), ni, nj, nk);
        hori_diff<1>("fused 32x8x8 TMA-staged", fused_t<32, 8, 8>(), ni, nj, nk);
        hori_diff<1>("fused 32x8x8 not staged", fused_t<32, 8, 8, 3, true, 4, true, 0, false>(), ni, nj, nk);
        hori_diff<1>("fused 32x8x4 TMA-staged", fused_t<32, 8, 4>(), ni, nj, nk);
        hori_diff<1>("fused 64x8x8 TMA-staged", fused_t<64, 8, 8>(), ni, nj, nk);
        // sweeps with L2 eviction priorities: flushed temporaries evict_last, streamed fields and dead temporaries evict_first
        vert_adv<1>("fused L2 hints, prefetch 2 L1", fused_t<32, 8, 8, 3, true, 2, true, 0, true, true, true>(), ni, nj, nk);
        vert_adv<1>("fused L2 hints, prefetch 4 L1", fused_t<32, 8, 8, 3, true, 4, true, 0, true, true, true>(), ni, nj, nk);
        vert_adv<1>("fused L2 hints, no prefetch", fused_t<32, 8, 8, 3, true, 0, true, 0, true, true, true>(), ni, nj, nk);
        vert_adv<1>("fused L2 hints, prefetch 4 L2", fused_t<32, 8, 8, 3, true, 4, false, 0, true, true, true>(), ni, nj, nk);
        vert_adv<1>("fused unroll 3 prefetch 4 L1", fused_t<32, 8, 8, 3, true, 4, true>(), ni, nj, nk);
        vert_adv<1>("fused unroll 3 prefetch 2 L1", fused_t<32, 8, 8, 3, true, 2, true>(), ni, nj, nk);
        vert_adv<1>("fused unroll 3 prefetch 3 L1", fused_t<32, 8, 8, 3, true, 3, true>(), ni, nj, nk);
        vert_adv<1>("fused unroll 4 prefetch 4 L1", fused_t<32, 8, 8, 4, true, 4, true>(), ni, nj, nk);
        vert_adv<1>("fused 32x4 unroll 3 prefetch 4", fused_t<32, 4, 8, 3, true, 4, true>(), ni, nj, nk);
#endif
    } catch (std::exception const &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    return 0;
}
