/*
 * gtb200/boundaries/boundary.hpp -- the reference's boundary-condition interface (boundaries/boundary.hpp:57-72) on
 * top of libgtb200.so for the predefined conditions.
 *
 *     #include <gtb200/boundaries/boundary.hpp>
 *     namespace bd = gtb200::boundaries;
 *     std::array<gtb_halo_desc, 3> halos = {{{2, 2, 2, ni + 1, ni_total}, {2, 2, 2, nj + 1, nj + 4}, {0, 0, 0, nk - 1, nk}}};
 *     bd::make_boundary(halos, bd::value_boundary<double>(3.5)).apply(a, b);      // a, b: double * to element (0,0,0)
 *     bd::make_boundary(halos, bd::copy_boundary(), pred).apply(dst, src);        // pred(ei, ej, ek) -> bool
 *
 * Mirrors gridtools::boundaries::boundary<BoundaryFunction, Arch, Predicate>: same constructor arguments (three halo
 * descriptors in increasing-stride order, the condition, a run-time predicate over the 26 directions,
 * predicate.hpp:24-30 / grid_predicate.hpp:20-34), same apply(fields...) with the fields' roles defined by the
 * condition (value.hpp:27-66: every field is assigned; copy.hpp:26-46: the last field is the source).  What differs:
 * the conditions are the predefined ones (a user functor would have to be compiled into a kernel: the generic path of
 * stencil/b200.hpp is the place for that), fields are raw device pointers to storage element (0,0,0) -- what
 * data_store::get_target_ptr() returns -- and the whole call is ONE launch on `stream` for all directions and fields
 * (apply_gpu.hpp:236-313 launches per-direction thread blocks over the bounding box of all halos).
 */
#pragma once

#include <array>
#include <stdexcept>
#include <string>

#include "../../gtb200.h"

namespace gtb200 {
    namespace boundaries {

        /// value.hpp:27-66 (T{} is zero_boundary, zero.hpp)
        template <class T>
        struct value_boundary {
            T value{};
            value_boundary() = default;
            explicit value_boundary(T const &v) : value(v) {}
            static constexpr int kind = GTB_BC_VALUE;
            double as_double() const { return static_cast<double>(value); }
            using element_type = T;
        };
        template <class T>
        using zero_boundary = value_boundary<T>;

        /// copy.hpp:26-46: every field but the last receives the last
        struct copy_boundary {
            static constexpr int kind = GTB_BC_COPY;
            double as_double() const { return 0; }
            using element_type = void;
        };

        /// predicate.hpp:24-30
        struct default_predicate {
            bool operator()(int, int, int) const { return true; }
        };

        template <class Condition, class Predicate = default_predicate>
        class boundary {
            std::array<gtb_halo_desc, 3> m_halos;
            Condition m_condition;
            int m_mask[27];

            template <class T>
            void run(T *const *fields, int n, void *stream) const {
                static_assert(sizeof(T) == 4 || sizeof(T) == 8, "float or double fields");
                void *ptrs[16];
                if (n > 16)
                    throw std::runtime_error("gtb200::boundaries::boundary: at most 16 fields per apply()");
                for (int f = 0; f < n; ++f)
                    ptrs[f] = const_cast<void *>(static_cast<const void *>(fields[f]));
                int st = gtb_boundary_apply(
                    m_halos.data(), m_mask, Condition::kind, m_condition.as_double(), ptrs, n, (int)sizeof(T), stream);
                if (st != GTB_OK)
                    throw std::runtime_error(std::string("gtb_boundary_apply: ") + gtb_last_error());
            }

          public:
            boundary(std::array<gtb_halo_desc, 3> const &halos, Condition const &condition, Predicate predicate = Predicate())
                : m_halos(halos), m_condition(condition) {
                for (int ek = -1; ek <= 1; ++ek)
                    for (int ej = -1; ej <= 1; ++ej)
                        for (int ei = -1; ei <= 1; ++ei)
                            m_mask[(ei + 1) + 3 * (ej + 1) + 9 * (ek + 1)] =
                                (ei || ej || ek) ? (predicate(ei, ej, ek) ? 1 : 0) : 0;
            }

            /// apply(fields...) on the legacy default stream, like the reference
            template <class T, class... Ts>
            void apply(T *first, Ts *...rest) const {
                T *fields[] = {first, rest...};
                run(fields, 1 + (int)sizeof...(rest), nullptr);
            }
            /// same on an explicit stream (cudaStream_t passed as void *)
            template <class T, class... Ts>
            void apply_on(void *stream, T *first, Ts *...rest) const {
                T *fields[] = {first, rest...};
                run(fields, 1 + (int)sizeof...(rest), stream);
            }
        };

        template <class Condition, class Predicate = default_predicate>
        boundary<Condition, Predicate> make_boundary(
            std::array<gtb_halo_desc, 3> const &halos, Condition const &condition, Predicate predicate = Predicate()) {
            return boundary<Condition, Predicate>(halos, condition, predicate);
        }
    } // namespace boundaries
} // namespace gtb200
