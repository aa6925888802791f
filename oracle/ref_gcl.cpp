/*
 * ref_gcl.cpp -- the UNMODIFIED reference gcl (gridtools/gcl/halo_exchange.hpp and everything below it:
 * high_level/descriptors.hpp:61-91,370-641, high_level/descriptor_generic_manual.hpp, low_level/Halo_Exchange_3D.hpp,
 * low_level/proc_grids_3D.hpp) compiled against oracle/mpi_shim/mpi.h and run with THREADS AS RANKS.
 *
 * TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libgtref.so): pins the restated index algebra of gt_oracle.c
 * (gto_halo_*), generates tests/golden/halo_*.npz, and is the expectation of the GPU exchange tests.  No reference
 * source is copied; the headers are #included where they lie.
 *
 * The driver mirrors what tests/regression/gcl/test_halo_exchange_3D.cpp:195-208,241-248 does per rank: build the
 * Cartesian communicator, construct the pattern, register the halos in USER dimension order, setup, pack,
 * exchange, unpack.
 */
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

#include <mpi.h> // oracle/mpi_shim/mpi.h

#include <gridtools/common/array.hpp>
#include <gridtools/common/halo_descriptor.hpp>
#include <gridtools/common/layout_map.hpp>
#include <gridtools/gcl/halo_exchange.hpp>

namespace {
    namespace gt = gridtools;
    namespace gcl = gridtools::gcl;

    struct elem16 {
        int v[4];
    };

    struct job {
        int layout[3];       // T_layout_map values (GridTools convention: 2 = unit stride)
        int proc_layout[3];  // layout2proc_map_abs
        int proc_dims[3];
        int periodic[3];     // USER dimension order (what the ctor's boollist takes)
        const int *halos;    // dynamic_ut: 15 ints (user dim order); generic: n_fields * 15
        int elem_size, n_fields;
        void **fields;       // [rank * n_fields + f]
        int use_vector, generic, split_phase;
    };

    template <class Layout, class ProcLayout, class T>
    void rank_body(job const &j, int rank) {
        MPI_Comm cart;
        int period[3] = {1, 1, 1};
        int dims[3] = {j.proc_dims[0], j.proc_dims[1], j.proc_dims[2]};
        MPI_Cart_create(MPI_COMM_WORLD, 3, dims, period, false, &cart);
        auto hd = [&](int f, int d) {
            const int *h = j.halos + (f * 3 + d) * 5;
            return gt::halo_descriptor(h[0], h[1], h[2], h[3], h[4]);
        };
        typename gcl::MPI_3D_process_grid_t<3>::period_type per(j.periodic[0], j.periodic[1], j.periodic[2]);
        std::vector<T *> ptrs(j.n_fields);
        for (int f = 0; f < j.n_fields; ++f)
            ptrs[f] = static_cast<T *>(j.fields[rank * j.n_fields + f]);
        auto run_exchange = [&](auto &testee) {
            if (j.split_phase) { // gcl/halo_exchange.hpp:286-304
                testee.post_receives();
                testee.do_sends();
                testee.wait();
            } else
                testee.exchange();
        };
        if (!j.generic) {
            gcl::halo_exchange_dynamic_ut<Layout, ProcLayout, T, gcl::cpu> testee(per, cart);
            testee.template add_halo<0>(hd(0, 0));
            testee.template add_halo<1>(hd(0, 1));
            testee.template add_halo<2>(hd(0, 2));
            testee.setup(j.n_fields);
            if (j.use_vector || j.n_fields != 3) {
                testee.pack(ptrs);
                run_exchange(testee);
                testee.unpack(ptrs);
            } else {
                testee.pack(ptrs[0], ptrs[1], ptrs[2]);
                run_exchange(testee);
                testee.unpack(ptrs[0], ptrs[1], ptrs[2]);
            }
        } else {
            using testee_t = gcl::halo_exchange_generic<ProcLayout, gcl::cpu>;
            testee_t testee(per, cart);
            // the enclosing halo: per dimension the maximum over the fields (test_halo_exchange_3D.cpp:224-236)
            gt::array<gt::halo_descriptor, 3> enclosing;
            for (int d = 0; d < 3; ++d) {
                int m = 0, p = 0, len = 0;
                for (int f = 0; f < j.n_fields; ++f) {
                    const int *h = j.halos + (f * 3 + d) * 5;
                    m = std::max(m, h[0]);
                    p = std::max(p, h[1]);
                    len = std::max(len, h[3] - h[2] + 1);
                }
                enclosing[d] = gt::halo_descriptor(m, p, m, len + m - 1, len + m + p);
            }
            testee.setup(j.n_fields,
                gcl::field_on_the_fly<int, Layout, testee_t::template traits>(nullptr, enclosing),
                sizeof(T));
            using fotf_t = gcl::field_on_the_fly<T, Layout, testee_t::template traits>;
            std::vector<fotf_t> fs;
            for (int f = 0; f < j.n_fields; ++f)
                fs.emplace_back(ptrs[f], gt::array<gt::halo_descriptor, 3>{hd(f, 0), hd(f, 1), hd(f, 2)});
            if (j.use_vector || j.n_fields != 3) {
                testee.pack(fs);
                run_exchange(testee);
                testee.unpack(fs);
            } else {
                testee.pack(fs[0], fs[1], fs[2]);
                run_exchange(testee);
                testee.unpack(fs[0], fs[1], fs[2]);
            }
        }
        MPI_Barrier(MPI_COMM_WORLD);
        MPI_Comm_free(&cart);
    }

    template <class Layout, class ProcLayout>
    int by_type(job const &j) {
        int n = j.proc_dims[0] * j.proc_dims[1] * j.proc_dims[2];
        switch (j.elem_size) {
        case 4:
            mpi_shim::run(n, [&](int r) { rank_body<Layout, ProcLayout, float>(j, r); });
            return 0;
        case 8:
            mpi_shim::run(n, [&](int r) { rank_body<Layout, ProcLayout, double>(j, r); });
            return 0;
        case 16:
            mpi_shim::run(n, [&](int r) { rank_body<Layout, ProcLayout, elem16>(j, r); });
            return 0;
        }
        return 2;
    }

    template <class Layout>
    int by_proc_layout(job const &j) {
        auto is = [&](int a, int b, int c) {
            return j.proc_layout[0] == a && j.proc_layout[1] == b && j.proc_layout[2] == c;
        };
        if (is(0, 1, 2))
            return by_type<Layout, gt::layout_map<0, 1, 2>>(j);
        if (is(1, 0, 2))
            return by_type<Layout, gt::layout_map<1, 0, 2>>(j);
        if (is(2, 1, 0))
            return by_type<Layout, gt::layout_map<2, 1, 0>>(j);
        if (is(1, 2, 0)) // not an involution: user dimension d -> process dimension P[d] differs from its inverse
            return by_type<Layout, gt::layout_map<1, 2, 0>>(j);
        if (is(2, 0, 1))
            return by_type<Layout, gt::layout_map<2, 0, 1>>(j);
        return 3;
    }

    int by_layout(job const &j) {
        auto is = [&](int a, int b, int c) { return j.layout[0] == a && j.layout[1] == b && j.layout[2] == c; };
        if (is(0, 1, 2))
            return by_proc_layout<gt::layout_map<0, 1, 2>>(j);
        if (is(0, 2, 1))
            return by_proc_layout<gt::layout_map<0, 2, 1>>(j);
        if (is(1, 0, 2))
            return by_proc_layout<gt::layout_map<1, 0, 2>>(j);
        if (is(1, 2, 0))
            return by_proc_layout<gt::layout_map<1, 2, 0>>(j);
        if (is(2, 0, 1))
            return by_proc_layout<gt::layout_map<2, 0, 1>>(j);
        if (is(2, 1, 0))
            return by_proc_layout<gt::layout_map<2, 1, 0>>(j);
        return 1;
    }
} // namespace

extern "C" {
#define GTREF_API __attribute__((visibility("default")))

/* One complete pack / exchange / unpack of the reference's halo_exchange_dynamic_ut<layout, proc_layout, T, cpu>
 * (generic = 0) or halo_exchange_generic<proc_layout, cpu> (generic = 1) on proc_dims[0]*[1]*[2] in-process ranks.
 * layout / proc_layout are the template arguments as the reference's user writes them; periodic and halos are in
 * USER dimension order; fields[r*n_fields+f] points at storage element (0,0,0) of field f on rank r.
 * Returns 0, or 1/2/3 for an unsupported layout / element size / proc layout. */
GTREF_API int gtref_gcl_exchange(const int layout[3], const int proc_layout[3], const int proc_dims[3],
    const int periodic[3], const int *halos, int elem_size, int n_fields, void **fields, int use_vector, int generic,
    int split_phase) {
    job j;
    for (int d = 0; d < 3; ++d) {
        j.layout[d] = layout[d];
        j.proc_layout[d] = proc_layout[d];
        j.proc_dims[d] = proc_dims[d];
        j.periodic[d] = periodic[d];
    }
    j.halos = halos;
    j.elem_size = elem_size;
    j.n_fields = n_fields;
    j.fields = fields;
    j.use_vector = use_vector;
    j.generic = generic;
    j.split_phase = split_phase;
    return by_layout(j);
}

/* MPI_3D_process_grid_t::proc(I,J,K) of the rank at `coords` (proc_grids_3D.hpp:179-211) through the shim. */
GTREF_API int gtref_gcl_proc(const int proc_dims[3], const int periodic[3], int rank, int di, int dj, int dk) {
    int n = proc_dims[0] * proc_dims[1] * proc_dims[2], res = -2;
    mpi_shim::run(n, [&](int r) {
        MPI_Comm cart;
        int period[3] = {1, 1, 1};
        int dims[3] = {proc_dims[0], proc_dims[1], proc_dims[2]};
        MPI_Cart_create(MPI_COMM_WORLD, 3, dims, period, false, &cart);
        {
            gcl::MPI_3D_process_grid_t<3> g(
                gcl::MPI_3D_process_grid_t<3>::period_type(periodic[0], periodic[1], periodic[2]), cart);
            if (r == rank)
                res = g.proc(di, dj, dk);
        }
        MPI_Comm_free(&cart);
    });
    return res;
}
}
