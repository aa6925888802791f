// b200_generic.cu -- the generic paths of the stencil::b200 tag (specs NOT bound to a hand-written kernel) on the device:
// every case of cases.hpp through GridTools' own frontend on storage::gpu stores, against the reference's cpu_ifirst
// backend on the same inputs in the same process (whole storages: the halo must come back untouched).
//   * st::b200<>                                  fused path with its defaults (include/gtb200/stencil/b200_fused.hpp)
//   * st::b200<stream, stage_by_stage>            the fallback (run_generic in b200.hpp)
//   * st::b200<stream, fused, block_geometry<...>> sweeps in separate launches, no prefetch, other unroll factor
// Built here against /root/reference/include; the binary travels to the GPU box.  Prints one line per case and
// "ALL PASSED" / "FAILED"; exit code 0 / 1.
#include <cstdio>
#include <string>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/cpu_ifirst.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/gpu.hpp>

#include <gtb200/stencil/b200.hpp>

#include "cases.hpp"

namespace {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;

    int g_failed = 0;

    template <class Backend>
    void test_generic_cases(std::string label, int ni, int nj, int nk) {
        gt::storage::gpu dev;
        gt::storage::cpu_ifirst host;
        st::cpu_ifirst<> ref_be;
        Backend be;
        auto fwd = [] { return st::execute_forward(); };
        auto bwd = [] { return st::execute_backward(); };
        auto name = [&](const char *what) {
            return label + " " + what + " " + std::to_string(ni) + "x" + std::to_string(nj) + "x" + std::to_string(nk);
        };
        cases::same(name("hori_diff f64").c_str(), cases::hori_diff<double>(dev, be, ni, nj, nk),
            cases::hori_diff<double>(host, ref_be, ni, nj, nk), ni + 4, nj + 4, nk, 1e-12, g_failed);
        cases::same(name("hori_diff f32").c_str(), cases::hori_diff<float>(dev, be, ni, nj, nk),
            cases::hori_diff<float>(host, ref_be, ni, nj, nk), ni + 4, nj + 4, nk, 1e-5, g_failed);
        cases::same(name("simple_hori_diff f64").c_str(), cases::simple_hori_diff<double>(dev, be, ni, nj, nk),
            cases::simple_hori_diff<double>(host, ref_be, ni, nj, nk), ni + 4, nj + 4, nk, 1e-12, g_failed);
        cases::same(name("vert_adv f64").c_str(), cases::vert_adv<double>(dev, be, ni, nj, nk),
            cases::vert_adv<double>(host, ref_be, ni, nj, nk), ni + 6, nj + 6, nk, 1e-12, g_failed);
        cases::same(name("vert_adv f32").c_str(), cases::vert_adv<float>(dev, be, ni, nj, nk),
            cases::vert_adv<float>(host, ref_be, ni, nj, nk), ni + 6, nj + 6, nk, 1e-4, g_failed);
        cases::same(name("tridiagonal").c_str(), cases::tridiagonal(dev, be, ni, nj, 6),
            cases::tridiagonal(host, ref_be, ni, nj, 6), ni, nj, 6, 1e-12, g_failed);
        cases::same(name("k-cache fill forward").c_str(), cases::kcache_fill(fwd, dev, be, ni, nj, nk),
            cases::kcache_fill(fwd, host, ref_be, ni, nj, nk), ni, nj, nk, 0, g_failed);
        cases::same(name("k-cache fill backward").c_str(), cases::kcache_fill(bwd, dev, be, ni, nj, nk),
            cases::kcache_fill(bwd, host, ref_be, ni, nj, nk), ni, nj, nk, 0, g_failed);
        for (bool forward : {true, false}) {
            cases::same(name(forward ? "k-cache flush forward" : "k-cache flush backward").c_str(),
                cases::kcache_flush(forward, dev, be, ni, nj, nk), cases::kcache_flush(forward, host, ref_be, ni, nj, nk),
                ni, nj, nk, 0, g_failed);
            cases::same(name(forward ? "k-cache fill+flush forward" : "k-cache fill+flush backward").c_str(),
                cases::kcache_fill_and_flush(forward, dev, be, ni, nj, nk),
                cases::kcache_fill_and_flush(forward, host, ref_be, ni, nj, nk), ni, nj, nk, 0, g_failed);
        }
        cases::same(name("k-cache local, two stages").c_str(), cases::kcache_local(dev, be, ni, nj, nk),
            cases::kcache_local(host, ref_be, ni, nj, nk), ni, nj, nk, 1e-14, g_failed);
        cases::same(name("mixed tiles + plain temporaries").c_str(), cases::mixed<double>(dev, be, ni, nj, 2, nk),
            cases::mixed<double>(host, ref_be, ni, nj, 2, nk), ni + 4, nj + 4, nk + 2, 1e-12, g_failed);
        {
            int bad = cases::prepare_tracers(dev, be, ni, nj, nk, 5);
            std::printf("%-58s %s (%d of 5 tracers differ)\n", name("expandable_run<2>, 5 tracers").c_str(),
                bad ? "FAILED" : "ok", bad);
            g_failed += bad != 0;
        }
        cases::same(name("forward sweep with IJ extents").c_str(), cases::sweep_with_extents(dev, be, ni, nj, nk),
            cases::sweep_with_extents(host, ref_be, ni, nj, nk), ni + 6, nj + 6, nk, 1e-12, g_failed);
        cases::same(name("temporary read at IJ offsets, not cached").c_str(),
            cases::mixed_plain<double>(dev, be, ni, nj, 2, nk), cases::mixed_plain<double>(host, ref_be, ni, nj, 2, nk),
            ni + 4, nj + 4, nk + 2, 1e-12, g_failed);
    }
} // namespace

int main() {
    try {
        using staged_t = st::b200<gtb200::default_stream, gtb200::stage_by_stage>;
        // sweeps in separate launches (every read-only field on the ld.global.nc path), no prefetch
        using unchained_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
            gtb200::block_geometry<32, 8, 8, 3, false, 0, false>>;
        test_generic_cases<st::b200<>>("fused", 70, 19, 13); // partial tiles in i and j, partial k block
        test_generic_cases<st::b200<>>("fused", 1, 1, 2);
        test_generic_cases<st::b200<>>("fused", 128, 64, 80);
        test_generic_cases<staged_t>("staged", 70, 19, 13);
        test_generic_cases<unchained_t>("fused, sweeps unchained", 70, 19, 13);
        // parallel multi-stages whose temporaries are all ij caches on per-thread register tiles
        using rtiles_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
            gtb200::block_geometry<32, 16, 4, 3, true, 4, true, 0, true, true>>;
        using rtiles_plain_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
            gtb200::block_geometry<64, 8, 2, 3, true, 4, true, 0, false, true>>;
        test_generic_cases<rtiles_t>("register tiles", 70, 19, 13);
        test_generic_cases<rtiles_t>("register tiles", 128, 64, 80);
        test_generic_cases<rtiles_plain_t>("register tiles, not staged", 70, 19, 13);
        // sweeps with L2 eviction priorities on their flushes, fills and streamed fields
        using hinted_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
            gtb200::block_geometry<32, 8, 4, 3, true, 4, true, 0, true, true, true>>;
        test_generic_cases<hinted_t>("L2 hints", 70, 19, 13);
        // ... and the shared-memory tile path for the same multi-stages
        using tiles_t = st::b200<gtb200::default_stream, gtb200::fused_when_possible,
            gtb200::block_geometry<32, 8, 4, 3, true, 4, true, 0, true, false>>;
        test_generic_cases<tiles_t>("shared-memory tiles", 70, 19, 13);
        test_generic_cases<tiles_t>("shared-memory tiles", 128, 64, 80);
    } catch (std::exception const &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::puts(g_failed ? "FAILED" : "ALL PASSED");
    return g_failed ? 1 : 0;
}
