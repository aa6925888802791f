// Driver for the reference's MPI tests (tests/regression/gcl/*.cpp) with THREADS AS RANKS: every rank thread runs
// all tests, like every MPI process does under the reference's tests/src/mpi_test_driver.cpp.  TEST INFRASTRUCTURE.
//   gcl_reference_<arch> [n_ranks] [--gtest_filter=...]
#include <cstdlib>
#include <atomic>

#include <mpi.h> // oracle/mpi_shim/mpi.h

#include <gtest/gtest.h>

#include <gridtools/gcl/GCL.hpp>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

int main(int argc, char **argv) {
    testing::InitGoogleTest(&argc, argv);
    int n = argc > 1 ? std::atoi(argv[1]) : 4;
    std::atomic<int> failed{0};
    testing::internal::state().quiet = true;
    testing::internal::state().between_tests = [] { MPI_Barrier(MPI_COMM_WORLD); };
    mpi_shim::run(n, [&](int rank) {
#ifdef __CUDACC__
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0)
            cudaSetDevice(rank % ndev);
#endif
        gridtools::gcl::init();
        failed += RUN_ALL_TESTS();
        MPI_Barrier(MPI_COMM_WORLD);
    });
    std::printf("%s (%d ranks)\n", failed ? "SOME FAILED" : "ALL PASSED", n);
    return failed ? 1 : 0;
}
