// emulated_cta.hpp -- TEST INFRASTRUCTURE: runs the per-thread body of the fused generic path
// (include/gtb200/stencil/b200_fused.hpp) on the host, one OpenMP team per CTA, `#pragma omp barrier` standing in for
// `__syncthreads` and a heap buffer for the CTA's shared memory.  It exists so that the thread-to-point mapping, the
// shared-memory tiles, the register windows and their fill / flush can be pinned against the reference's cpu_ifirst
// backend in the build container, which has no GPU.  Nothing under include/ or gridtools_b200/ refers to it.
#pragma once

#include <memory>
#include <vector>

#include <omp.h>

#include <gridtools/sid/allocator.hpp>
#include <gtb200/stencil/b200_fused.hpp>

namespace emulated {
    using gridtools::int_t;

    struct state {
        int_t block[3];
        char *smem;
    };
    inline state &cta_state() {
        static state s{};
        return s;
    }

    struct cta {
        static int_t tid() { return omp_get_thread_num(); }
        static int_t block_i() { return cta_state().block[0]; }
        static int_t block_j() { return cta_state().block[1]; }
        static int_t block_k() { return cta_state().block[2]; }
        static void sync() {
#pragma omp barrier
        }
        static char *smem() { return cta_state().smem; }
    };

    struct launcher {
        using cta_t = cta;
        long launches = 0, syncs_possible = 0;

        static auto allocator() { return gridtools::sid::cached_allocator(&std::make_unique<char[]>); }
        // no TMA on the host: the staged tiles are filled by the cooperative copy
        static bool tensor_map(void *, const void *, int, const int64_t *, const int64_t *, const int *) {
            ++staged_fields();
            return false;
        }
        static long &staged_fields() { // how many fields have been staged through emulated shared memory so far
            static long n = 0;
            return n;
        }

        static long &register_tile_launches() { // launches of multi-stages that ran on per-thread register tiles
            static long n = 0;
            return n;
        }
        template <class Body, class = void>
        struct on_register_tiles : std::false_type {};
        template <class Body>
        struct on_register_tiles<Body, std::void_t<typename Body::register_tiles_t>> : std::true_type {};

        template <class Body>
        void launch(Body const &body, int_t nbi, int_t nbj, int_t nbk, int_t threads, int_t smem) {
            ++launches;
            register_tile_launches() += on_register_tiles<Body>::value;
            std::vector<char> shared(smem + 16);
            omp_set_dynamic(0);
            for (int_t bk = 0; bk < nbk; ++bk)
                for (int_t bj = 0; bj < nbj; ++bj)
                    for (int_t bi = 0; bi < nbi; ++bi) {
                        cta_state() = {{bi, bj, bk}, shared.data()};
                        // poison the tiles: nothing may be read that this CTA did not write
                        for (auto &c : shared)
                            c = char(0xff);
#pragma omp parallel num_threads(threads)
                        {
                            if (omp_get_num_threads() != threads)
                                std::abort();
                            body();
                        }
                    }
        }
    };

    /// backend tag: the fused path of stencil::b200, executed by the emulated CTAs
    template <class Geo = gridtools::stencil::b200_backend::fused::geometry<>>
    struct backend {
        template <class Spec, class Grid, class DataStores>
        friend void gridtools_backend_entry_point(backend, Spec spec, Grid const &grid, DataStores data_stores) {
            launcher l;
            gridtools::stencil::b200_backend::fused::run_fused_spec<Geo>(l, spec, grid, std::move(data_stores));
            last_launches() = l.launches;
        }
        static long &last_launches() {
            static long n = 0;
            return n;
        }
    };
} // namespace emulated
