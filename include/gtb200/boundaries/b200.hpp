/*
 * gtb200/boundaries/b200.hpp -- boundary conditions for the arch tag gridtools::gcl::b200, behind the reference's own
 * class templates:
 *
 *     #include <gridtools/boundaries/boundary.hpp>
 *     #include <gtb200/boundaries/b200.hpp>
 *     gridtools::boundaries::boundary<my_functor, gridtools::gcl::b200, predicate_t>(halos, my_functor{...}, pred)
 *         .apply(field_a, field_b);                                        // was: gcl::gpu
 *     gridtools::boundaries::distributed_boundaries<comm_traits<storage_t, gridtools::gcl::b200, timer_cuda>> ...
 *
 * boundaries/boundary.hpp:31-44 selects the implementation by arch tag (_impl::select_apply); this header adds the
 * case for gcl::b200.  Any boundary functor works -- the predefined value_boundary / zero_boundary / copy_boundary
 * (value.hpp, zero.hpp, copy.hpp) and user functors `operator()(direction<I, J, K>, views..., i, j, k)` alike; the
 * functor is instantiated in a kernel in the user's translation unit (nvcc), like in the reference's GPU path
 * (apply_gpu.hpp:236-313).  With it distributed_boundaries<comm_traits<Storage, gcl::b200, Timer>>
 * (distributed_boundaries.hpp:68,161) works unchanged: its pattern type is
 * gcl::halo_exchange_dynamic_ut<..., gcl::b200> (gtb200/gcl/b200.hpp) and its boundary pass is this one.
 *
 * What differs from apply_gpu.hpp: the 26 outside regions are walked as flat element lists by ONE small grid with a
 * grid stride (the reference launches a 3-d grid over the bounding box of the largest region, most threads of which
 * are idle for every other region), on an optional stream.
 */
#pragma once

#include <gridtools/boundaries/boundary.hpp>
#include <gridtools/boundaries/direction.hpp>
#include <gridtools/boundaries/predicate.hpp>
#include <gridtools/common/array.hpp>
#include <gridtools/common/cuda_util.hpp>
#include <gridtools/common/halo_descriptor.hpp>
#include <gridtools/common/host_device.hpp>

#include "../gcl/arch.hpp"

namespace gridtools {
    namespace boundaries {
        namespace b200_impl_ {
            struct region {
                int lo[3], len[3];
                long long count; // 0: direction not selected by the predicate, or empty
            };
            struct regions {
                region r[27]; // n = (I + 1) * 9 + (J + 1) * 3 + (K + 1)
            };

            template <int N>
            struct walk {
                template <class BF, class... Views>
                static GT_FUNCTION_DEVICE void run(BF const &bf, regions const &table, long long tid, long long nthreads,
                    Views const &...views) {
                    constexpr int I = N / 9 - 1, J = (N / 3) % 3 - 1, K = N % 3 - 1;
                    region const &r = table.r[N];
                    for (long long e = tid; e < r.count; e += nthreads) {
                        const long long q = e / r.len[0];
                        const int i = int(e - q * r.len[0]), j = int(q % r.len[1]), k = int(q / r.len[1]);
                        bf(direction<sign(I), sign(J), sign(K)>(), views..., uint_t(r.lo[0] + i), uint_t(r.lo[1] + j),
                            uint_t(r.lo[2] + k));
                    }
                    walk<N + 1>::run(bf, table, tid, nthreads, views...);
                }
            };
            template <>
            struct walk<13> { // the centre is no boundary
                template <class BF, class... Views>
                static GT_FUNCTION_DEVICE void run(BF const &bf, regions const &table, long long tid, long long nthreads,
                    Views const &...views) {
                    walk<14>::run(bf, table, tid, nthreads, views...);
                }
            };
            template <>
            struct walk<27> {
                template <class BF, class... Views>
                static GT_FUNCTION_DEVICE void run(BF const &, regions const &, long long, long long, Views const &...) {}
            };

#ifdef GT_CUDACC
            template <class BF, class... Views>
            __global__ void __launch_bounds__(256) bc_functor_kernel(BF const bf, regions const table, Views const... views) {
                walk<0>::run(bf, table, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x,
                    views...);
            }
#endif

            template <int N, class Predicate>
            void fill(regions &t, array<halo_descriptor, 3> const &hd, Predicate const &pred) {
                if constexpr (N < 27) {
                    constexpr int I = N / 9 - 1, J = (N / 3) % 3 - 1, K = N % 3 - 1;
                    region &r = t.r[N];
                    r.count = 0;
                    if (N != 13 && pred(direction<sign(I), sign(J), sign(K)>())) {
                        const int e[3] = {I, J, K};
                        r.count = 1;
                        for (int d = 0; d < 3; ++d) { // apply.hpp:44-56: loop_{low,high}_bound_outside of every dimension
                            r.lo[d] = hd[d].loop_low_bound_outside(e[d]);
                            r.len[d] = hd[d].loop_high_bound_outside(e[d]) - r.lo[d] + 1;
                            r.count *= r.len[d] > 0 ? r.len[d] : 0;
                        }
                    }
                    fill<N + 1>(t, hd, pred);
                }
            }
        } // namespace b200_impl_

        /// boundary_apply for gcl::b200 (the counterpart of boundary_apply_gpu, apply_gpu.hpp:236-313)
        template <class BoundaryFunction, class Predicate = default_predicate>
        struct boundary_apply_b200 {
          private:
            b200_impl_::regions m_regions;
            BoundaryFunction const m_boundary_function;
            void *m_stream = nullptr;
            long long m_total = 0;

          public:
            boundary_apply_b200(array<halo_descriptor, 3> const &hd, BoundaryFunction const &bf, Predicate predicate = Predicate())
                : m_boundary_function(bf) {
                b200_impl_::fill<0>(m_regions, hd, predicate);
                for (auto const &r : m_regions.r)
                    m_total += r.count;
            }
            boundary_apply_b200(array<halo_descriptor, 3> const &hd, Predicate predicate = Predicate())
                : boundary_apply_b200(hd, BoundaryFunction(), predicate) {}

            /// kernels are enqueued on this cudaStream_t (default: the legacy default stream, like the reference)
            void set_stream(void *cuda_stream) { m_stream = cuda_stream; }

            template <class... DataFieldViews>
            void apply(DataFieldViews const &...data_field_views) const {
#ifdef GT_CUDACC
                if (m_total == 0)
                    return;
                long long blocks = (m_total + 256 * 8 - 1) / (256 * 8);
                blocks = blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks); // at most four small blocks per SM
                b200_impl_::bc_functor_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(m_stream)>>>(
                    m_boundary_function, m_regions, data_field_views...);
                GT_CUDA_CHECK(cudaGetLastError());
#else
                static_assert(sizeof...(DataFieldViews) < 0,
                    "boundaries for gcl::b200 instantiate the boundary functor in a CUDA kernel: compile with nvcc");
#endif
            }
        };

        namespace _impl {
            template <class BoundaryFunction, class Predicate>
            struct select_apply<gcl::b200, BoundaryFunction, Predicate> {
                using type = boundary_apply_b200<BoundaryFunction, Predicate>;
            };
        } // namespace _impl
    } // namespace boundaries
} // namespace gridtools
