// fused_emulation.cpp -- TEST: the per-thread body of stencil::b200's fused generic path
// (include/gtb200/stencil/b200_fused.hpp), executed on the host by emulated CTAs (emulated_cta.hpp), against the
// reference's cpu_ifirst backend on the same inputs.  Plain g++ (-fopenmp), no GPU, no CUDA headers:
//
//   g++ -std=c++17 -O1 -fopenmp -I/root/reference/include -I include tests/cpp/fused_emulation.cpp
//
// prints one line per case and "ALL PASSED" / "FAILED"; exit code 0 / 1.
#include <cstdio>
#include <string>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/cpu_ifirst.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/cpu_kfirst.hpp>

#include "cases.hpp"
#include "emulated_cta.hpp"

namespace {
    namespace gt = gridtools;
    namespace st = gridtools::stencil;
    namespace fused = gridtools::stencil::b200_backend::fused;

    int g_failed = 0;

    template <class Geo>
    struct run {
        static constexpr long sweep_launches = Geo::chain_sweeps ? 1 : 2;
        using be_t = emulated::backend<Geo>;
        using ref_t = st::cpu_ifirst<>;
        // cpu_ifirst storage: i has the compile-time stride 1 of storage::gpu (what the staging of read-only fields
        // through shared memory needs, and padded rows)
        using traits_t = gt::storage::cpu_ifirst;
        std::string m_geo;

        void expect_launches(const char *what, long n) {
            if (be_t::last_launches() != n) {
                std::printf("%-58s FAILED (%ld launches, expected %ld)\n", what, be_t::last_launches(), n);
                ++g_failed;
            }
        }
        std::string name(std::string what, int ni, int nj, int nk) {
            return what + " " + std::to_string(ni) + "x" + std::to_string(nj) + "x" + std::to_string(nk) + " " + m_geo;
        }

        void all(int ni, int nj, int nk) {
            auto fwd = [] { return st::execute_forward(); };
            auto bwd = [] { return st::execute_backward(); };
            traits_t tr;
            {
                auto got = cases::hori_diff<double>(tr, be_t(), ni, nj, nk);
                expect_launches("hori_diff launches", 1);
                auto ref = cases::hori_diff<double>(tr, ref_t(), ni, nj, nk);
                cases::same(name("hori_diff f64", ni, nj, nk).c_str(), got, ref, ni + 4, nj + 4, nk, 1e-13, g_failed);
            }
            {
                // the one-stage call<> formulation on the emulated CTAs against the four-stage spec on cpu_ifirst
                auto got = cases::hori_diff_fused<double>(tr, be_t(), ni, nj, nk);
                expect_launches("hori_diff_fused launches", 1);
                auto ref = cases::hori_diff<double>(tr, ref_t(), ni, nj, nk);
                cases::same(name("hori_diff as one stage with call<> f64", ni, nj, nk).c_str(), got, ref, ni + 4, nj + 4,
                    nk, 1e-13, g_failed);
            }
            {
                auto got = cases::hori_diff<float>(tr, be_t(), ni, nj, 3);
                auto ref = cases::hori_diff<float>(tr, ref_t(), ni, nj, 3);
                cases::same(name("hori_diff f32", ni, nj, 3).c_str(), got, ref, ni + 4, nj + 4, 3, 1e-5, g_failed);
            }
            {
                auto got = cases::simple_hori_diff<double>(tr, be_t(), ni, nj, nk);
                expect_launches("simple_hori_diff launches", 1);
                auto ref = cases::simple_hori_diff<double>(tr, ref_t(), ni, nj, nk);
                cases::same(name("simple_hori_diff f64", ni, nj, nk).c_str(), got, ref, ni + 4, nj + 4, nk, 1e-13, g_failed);
            }
            {
                auto got = cases::vert_adv<double>(tr, be_t(), ni, nj, nk + 4);
                expect_launches("vert_adv launches", sweep_launches); // forward and backward sweeps chained
                auto ref = cases::vert_adv<double>(tr, ref_t(), ni, nj, nk + 4);
                cases::same(name("vert_adv f64", ni, nj, nk + 4).c_str(), got, ref, ni + 6, nj + 6, nk + 4, 1e-12, g_failed);
            }
            {
                auto got = cases::tridiagonal(tr, be_t(), ni, nj, 6);
                expect_launches("tridiagonal launches", sweep_launches);
                auto ref = cases::tridiagonal(tr, ref_t(), ni, nj, 6);
                cases::same(name("tridiagonal", ni, nj, 6).c_str(), got, ref, ni, nj, 6, 1e-13, g_failed);
            }
            {
                auto got = cases::kcache_fill(fwd, tr, be_t(), ni, nj, nk);
                auto ref = cases::kcache_fill(fwd, tr, ref_t(), ni, nj, nk);
                cases::same(name("k-cache fill forward", ni, nj, nk).c_str(), got, ref, ni, nj, nk, 0, g_failed);
                got = cases::kcache_fill(bwd, tr, be_t(), ni, nj, nk);
                ref = cases::kcache_fill(bwd, tr, ref_t(), ni, nj, nk);
                cases::same(name("k-cache fill backward", ni, nj, nk).c_str(), got, ref, ni, nj, nk, 0, g_failed);
            }
            for (bool forward : {true, false}) {
                auto got = cases::kcache_flush(forward, tr, be_t(), ni, nj, nk);
                auto ref = cases::kcache_flush(forward, tr, ref_t(), ni, nj, nk);
                cases::same(name(forward ? "k-cache flush forward" : "k-cache flush backward", ni, nj, nk).c_str(), got,
                    ref, ni, nj, nk, 0, g_failed);
                got = cases::kcache_fill_and_flush(forward, tr, be_t(), ni, nj, nk);
                ref = cases::kcache_fill_and_flush(forward, tr, ref_t(), ni, nj, nk);
                cases::same(
                    name(forward ? "k-cache fill+flush forward" : "k-cache fill+flush backward", ni, nj, nk).c_str(),
                    got, ref, ni, nj, nk, 0, g_failed);
            }
            {
                auto got = cases::kcache_local(tr, be_t(), ni, nj, nk);
                expect_launches("k-cache local launches", 1);
                auto ref = cases::kcache_local(tr, ref_t(), ni, nj, nk);
                cases::same(name("k-cache local, two stages", ni, nj, nk).c_str(), got, ref, ni, nj, nk, 0, g_failed);
            }
            {
                auto got = cases::mixed<double>(tr, be_t(), ni, nj, 2, nk);
                expect_launches("mixed launches", 2);
                auto ref = cases::mixed<double>(tr, ref_t(), ni, nj, 2, nk);
                cases::same(name("mixed tiles + plain temporaries, 2 intervals", ni, nj, nk + 2).c_str(), got, ref,
                    ni + 4, nj + 4, nk + 2, 1e-13, g_failed);
            }
            {
                auto got = cases::positional_sum(tr, be_t(), ni, nj, nk);
                auto ref = cases::positional_sum(tr, ref_t(), ni, nj, nk);
                cases::same(name("positional<i,j,k>", ni, nj, nk).c_str(), got, ref, ni + 2, nj + 2, nk, 0, g_failed);
            }
            {
                auto got = cases::parallel_multistage(tr, be_t(), ni, nj, nk);
                expect_launches("parallel multistage launches", 2);
                auto ref = cases::parallel_multistage(tr, ref_t(), ni, nj, nk);
                cases::same(name("two parallel multi-stages, temporary read at k+1", ni, nj, nk).c_str(), got, ref, ni,
                    nj, nk, 0, g_failed);
            }
            {
                auto got = cases::whole_axis(tr, be_t(), ni, nj, nk);
                auto ref = cases::whole_axis(tr, ref_t(), ni, nj, nk);
                cases::same(name("whole axis through a renamed k dimension", ni, nj, nk).c_str(), got, ref, ni, nj, nk, 0,
                    g_failed);
            }
            {
                auto got = cases::multi_types(tr, be_t(), ni, nj, nk);
                expect_launches("multi types launches", 1);
                auto ref = cases::multi_types(tr, ref_t(), ni, nj, nk);
                cases::same(name("float and double tiles, int field", ni, nj, nk).c_str(), got, ref, ni + 6, nj + 6, nk,
                    0, g_failed);
            }
            {
                int bad = cases::prepare_tracers(tr, be_t(), ni, nj, nk, 5);
                std::printf("%-58s %s (%d of 5 tracers differ)\n", name("expandable_run<2>, 5 tracers", ni, nj, nk).c_str(),
                    bad ? "FAILED" : "ok", bad);
                g_failed += bad != 0;
            }
            {
                auto got = cases::sweep_with_extents(tr, be_t(), ni, nj, nk);
                expect_launches("sweep with IJ extents launches", 2);
                auto ref = cases::sweep_with_extents(tr, ref_t(), ni, nj, nk);
                cases::same(name("forward sweep with IJ extents, flushed to a blocked temporary", ni, nj, nk).c_str(), got,
                    ref, ni + 6, nj + 6, nk, 1e-13, g_failed);
            }
            {
                auto got = cases::mixed_plain<double>(tr, be_t(), ni, nj, 2, nk);
                expect_launches("mixed (no caches) launches", 2);
                auto ref = cases::mixed_plain<double>(tr, ref_t(), ni, nj, 2, nk);
                cases::same(name("temporary read at IJ offsets, not cached", ni, nj, nk + 2).c_str(), got, ref, ni + 4,
                    nj + 4, nk + 2, 1e-13, g_failed);
            }
        }
    };
} // namespace

// c_array_copy.cpp:22-46: plain C arrays are SIDs too (compile-time strides, k fastest)
void c_arrays() {
    int out[7][5][3];
    decltype(out) in;
    int n = 0;
    for (auto &&vvv : in)
        for (auto &&vv : vvv)
            for (auto &&v : vv)
                v = n++;
    for (auto &&vvv : out)
        for (auto &&vv : vvv)
            for (auto &&v : vv)
                v = -1;
    st::run_single_stage(user::copy_f<1>(), emulated::backend<fused::geometry<4, 2, 2>>(), st::make_grid(7, 5, 3), in, out);
    int bad = 0;
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 5; ++j)
            for (int k = 0; k < 3; ++k)
                bad += out[i][j][k] != in[i][j][k];
    std::printf("%-58s %s (mismatches %d)\n", "copy between C arrays 7x5x3 [4x2x2 blocks]", bad ? "FAILED" : "ok", bad);
    g_failed += bad != 0;
}

// shared-memory tiles for the ij caches (RegisterTiles = false) ...
template <int BI = 32, int BJ = 8, int KB = 4, int U = 3, bool Chain = true, int P = 4, bool L1 = true, int PP = 0, bool Stage = true>
using tiles = fused::geometry<BI, BJ, KB, U, Chain, P, L1, PP, Stage, false>;
// ... or per-thread register tiles wherever a parallel multi-stage allows it (the default)
template <int BI = 32, int BJ = 8, int KB = 4, bool Stage = true>
using rtiles = fused::geometry<BI, BJ, KB, 3, true, 4, true, 0, Stage, true>;

int main() {
#ifndef GTB_ONLY_REGISTER_TILES
    c_arrays();
    const long before = emulated::launcher::register_tile_launches();
    // small blocks: many CTAs, partial tiles in i and j, partial k blocks
    run<tiles<8, 4, 3>>{"[8x4x3 blocks]"}.all(19, 9, 7);
    run<tiles<8, 4, 3>>{"[8x4x3 blocks]"}.all(8, 4, 3);
    run<tiles<8, 4, 3>>{"[8x4x3 blocks]"}.all(1, 1, 2);
    // the former default geometry on a domain of a few blocks
    run<tiles<32, 8, 8>>{"[32x8x8 blocks]"}.all(37, 11, 10);
    // extremes: one level per CTA with tiny blocks; all levels in one CTA with a wide block
    run<tiles<4, 2, 1>>{"[4x2x1 blocks]"}.all(9, 5, 4);
    run<tiles<64, 4, 80>>{"[64x4x80 blocks]"}.all(70, 6, 5);
    // prefetch ahead (a no-op on the host, but the address arithmetic is instantiated)
    run<tiles<8, 4, 3, 2, true, 4, true, 1>>{"[8x4x3 blocks, prefetch]"}.all(19, 9, 7);
    // sweeps in separate launches
    run<tiles<8, 4, 3, 2, false>>{"[8x4x3 blocks, unchained]"}.all(19, 9, 7);
    if (emulated::launcher::register_tile_launches() != before)
        ++g_failed; // none of the `tiles` configurations may have taken the register-tile path
#endif
    // parallel multi-stages whose temporaries are all ij caches on per-thread register tiles (the others as before)
    run<rtiles<8, 4, 3>>{"[8x4x3 blocks, register tiles]"}.all(19, 9, 7);
    run<rtiles<8, 4, 3>>{"[8x4x3 blocks, register tiles]"}.all(1, 1, 2);
    run<fused::geometry<>>{"[default geometry]"}.all(37, 11, 10);
    run<rtiles<4, 2, 1, false>>{"[4x2x1 blocks, register tiles, not staged]"}.all(9, 5, 4);
    std::printf("launches on register tiles: %ld\n", emulated::launcher::register_tile_launches());
    if (emulated::launcher::register_tile_launches() == 0)
        ++g_failed; // horizontal diffusion and friends must take that path in the register-tile configurations
    std::printf("fields staged through (emulated) shared memory: %ld\n", emulated::launcher::staged_fields());
    if (emulated::launcher::staged_fields() == 0)
        ++g_failed; // the staging of read-only fields of parallel multi-stages must be exercised here
    std::printf(g_failed ? "FAILED (%d)\n" : "ALL PASSED\n", g_failed);
    return g_failed ? 1 : 0;
}
