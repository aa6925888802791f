/*
 * gtb200/stencil/b200_shapes.hpp -- which specs the hand-written kernels of libgtb200.so may execute.
 *
 * A named kernel hard-codes a whole spec: number and execution of the multi-stages, which placeholder every stage
 * argument is wired to, the caches, the k intervals of every functor overload, the extents, the order of the run()
 * arguments.  Binding a spec to a kernel by its functor types alone (round 1) would silently run the hard-coded
 * stencil for a user spec that merely re-uses those functors with another wiring.  Here the binding is checked
 * against the frontend itself: for every kernel this header states the reference spec it implements, as a function
 * of the user's functor types (`reference_comp`), lets the reference's own frontend turn it into a backend spec
 * (core::convert_fe_to_be_spec, the very transformation stencil::run applies, stencil/core/backend.hpp:39), and
 * compares the result with the spec the backend was handed -- type for type, after the temporaries of both have been
 * renumbered in order of appearance (GT_DECLARE_TMP numbers them with __COUNTER__).  A spec that differs in ANY
 * respect takes the generic path.
 *
 * The reference specs are the ones of the reference's regression tests:
 *   copy              tests/regression/copy_stencil.cpp:42-47              run_single_stage(F, be, grid, in, out)
 *   hori_diff         tests/regression/horizontal_diffusion.cpp:98-112     4 stages, lap / flx / fly ij-cached
 *   hori_diff_fused   tests/regression/horizontal_diffusion_fused.cpp:92   run_single_stage(F, be, grid, out, in, coeff)
 *   simple_hori_diff  tests/regression/simple_hori_diff.cpp:73-80          2 stages, lap ij-cached, j-only coefficients
 *   vert_adv          tests/regression/vertical_advection_dycore.cpp:140-149  forward + backward, k caches
 *   tridiagonal       tests/regression/tridiagonal.cpp:83-87               forward_thomas + backward_thomas
 *   prepare_tracers   tests/regression/advection_pdbott_prepare_tracers.cpp:45-52  one stage (out, in, rho) of
 *                     expandable_run<Factor>: the backend sees chunks of Factor (and 1) copies of the stage
 */
#pragma once

#include <cstddef>
#include <type_traits>

#include <gridtools/meta.hpp>
#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/core/convert_fe_to_be_spec.hpp>
#include <gridtools/stencil/frontend/expandable_run.hpp>

namespace gridtools {
    namespace stencil {
        namespace b200_backend {
            namespace shape {
                // ------------------------------------------------------------ temporaries in order of appearance
                template <class T>
                struct collect_tmps {
                    using type = meta::list<>;
                };
                template <template <class...> class L, class... Ts>
                struct collect_tmps<L<Ts...>> {
                    using type = meta::concat<meta::list<>, typename collect_tmps<Ts>::type...>;
                };
                template <size_t N, class T>
                struct collect_tmps<cartesian::tmp_arg<N, T>> {
                    using type = meta::list<cartesian::tmp_arg<N, T>>;
                };

                constexpr size_t canonical_base = 1u << 20;

                template <class T, class Tmps>
                struct renumber {
                    using type = T;
                };
                template <template <class...> class L, class... Ts, class Tmps>
                struct renumber<L<Ts...>, Tmps> {
                    using type = L<typename renumber<Ts, Tmps>::type...>;
                };
                template <size_t N, class T, class Tmps>
                struct renumber<cartesian::tmp_arg<N, T>, Tmps> {
                    using type =
                        cartesian::tmp_arg<canonical_base + meta::st_position<Tmps, cartesian::tmp_arg<N, T>>::value, T>;
                };

                /// the spec with its temporaries numbered canonical_base, canonical_base + 1, ... in order of appearance
                template <class Spec>
                using normalized = typename renumber<Spec, meta::dedup<typename collect_tmps<Spec>::type>>::type;

                // ------------------------------------------------------------ the reference specs, by functor types
                enum class id { none, copy, hori_diff, hori_diff_fused, simple_hori_diff, vert_adv, tridiagonal, prepare_tracers };

                template <id K, class Functors, class T>
                struct reference_comp;

                template <class F, class T>
                struct reference_comp<id::copy, meta::list<F>, T> {
                    static constexpr size_t n_args = 2;
                    template <class... Args>
                    auto operator()(Args... args) const {
                        return execute_parallel().stage(F(), args...);
                    }
                };
                template <class F, class T>
                struct reference_comp<id::hori_diff_fused, meta::list<F>, T> {
                    static constexpr size_t n_args = 3;
                    template <class... Args>
                    auto operator()(Args... args) const {
                        return execute_parallel().stage(F(), args...);
                    }
                };
                template <class Lap, class Flx, class Fly, class Out, class T>
                struct reference_comp<id::hori_diff, meta::list<Lap, Flx, Fly, Out>, T> {
                    static constexpr size_t n_args = 3;
                    template <class In, class Coeff, class O>
                    auto operator()(In in, Coeff coeff, O out) const {
                        GT_DECLARE_TMP(T, lap, flx, fly);
                        return execute_parallel()
                            .ij_cached(lap, flx, fly)
                            .stage(Lap(), lap, in)
                            .stage(Flx(), flx, in, lap)
                            .stage(Fly(), fly, in, lap)
                            .stage(Out(), out, in, flx, fly, coeff);
                    }
                };
                template <class Wlap, class Divflux, class T>
                struct reference_comp<id::simple_hori_diff, meta::list<Wlap, Divflux>, T> {
                    static constexpr size_t n_args = 5;
                    template <class Coeff, class In, class O, class Cro, class Cru>
                    auto operator()(Coeff coeff, In in, O out, Cro crlato, Cru crlatu) const {
                        GT_DECLARE_TMP(T, lap);
                        return execute_parallel()
                            .ij_cached(lap)
                            .stage(Wlap(), lap, in, crlato, crlatu)
                            .stage(Divflux(), out, in, lap, crlato, coeff);
                    }
                };
                template <class Fwd, class Bwd, class T>
                struct reference_comp<id::vert_adv, meta::list<Fwd, Bwd>, T> {
                    static constexpr size_t n_args = 6;
                    template <class Us, class Ust, class W, class Up, class Ut, class Dtr>
                    auto operator()(Us utens_stage, Ust u_stage, W wcon, Up u_pos, Ut utens, Dtr dtr_stage) const {
                        GT_DECLARE_TMP(T, ccol, dcol, data_col);
                        return multi_pass(
                            execute_forward()
                                .k_cached(cache_io_policy::flush(), ccol, dcol)
                                .k_cached(cache_io_policy::fill(), u_stage)
                                .stage(Fwd(), utens_stage, wcon, u_stage, u_pos, utens, dtr_stage, ccol, dcol),
                            execute_backward().k_cached(data_col).stage(
                                Bwd(), utens_stage, u_pos, dtr_stage, ccol, dcol, data_col));
                    }
                };
                template <class Fwd, class Bwd, class T>
                struct reference_comp<id::tridiagonal, meta::list<Fwd, Bwd>, T> {
                    static constexpr size_t n_args = 5;
                    template <class Inf, class Diag, class Sup, class Rhs, class O>
                    auto operator()(Inf inf, Diag diag, Sup sup, Rhs rhs, O out) const {
                        return multi_pass(execute_forward().stage(Fwd(), inf, diag, sup, rhs),
                            execute_backward().stage(Bwd(), out, sup, rhs));
                    }
                };

                // ------------------------------------------------------------ does the backend's spec have that shape?
                template <class Comp, class Interval, class DataStores, class Indices>
                struct expected_spec;
                template <class Comp, class Interval, class DataStores, size_t... Is>
                struct expected_spec<Comp, Interval, DataStores, std::index_sequence<Is...>> {
                    using fe_spec_t = decltype(std::declval<Comp>()(frontend_impl_::arg<Is>()...));
                    using type = core::convert_fe_to_be_spec<fe_spec_t, Interval, DataStores>;
                };

                template <class DataStores, size_t... Is>
                constexpr bool keys_are_args(std::index_sequence<Is...>) {
                    return std::is_same<get_keys<DataStores>, hymap::keys<frontend_impl_::arg<Is>...>>::value;
                }

                /// true iff `Spec` (what gridtools_backend_entry_point received for run(comp, backend, grid, fields...))
                /// is exactly the reference spec of kernel K built from the functor list `Functors`.
                template <id K, class Functors, class T, class Spec, class Grid, class DataStores, class = void>
                struct matches : std::false_type {};

                template <id K, class Functors, class T, class Spec, class Grid, class DataStores>
                struct matches<K,
                    Functors,
                    T,
                    Spec,
                    Grid,
                    DataStores,
                    std::enable_if_t<(sizeof(reference_comp<K, Functors, T>) > 0) &&
                                     tuple_util::size<DataStores>::value == reference_comp<K, Functors, T>::n_args &&
                                     keys_are_args<DataStores>(
                                         std::make_index_sequence<reference_comp<K, Functors, T>::n_args>())>> {
                    using comp_t = reference_comp<K, Functors, T>;
                    using expected_t = typename expected_spec<comp_t,
                        typename Grid::interval_t,
                        DataStores,
                        std::make_index_sequence<comp_t::n_args>>::type;
                    static constexpr bool value = std::is_same<normalized<Spec>, normalized<expected_t>>::value;
                };

                // prepare_tracers: a chunk of `Factor` tracers of expandable_run (frontend/expandable_run.hpp:137-141):
                // data stores (out_0 .. out_F-1, in_0 .. in_F-1, rho) keyed expanded<J, arg<I>>
                template <class F, class T, class Spec, class Grid, class DataStores>
                struct matches<id::prepare_tracers,
                    meta::list<F>,
                    T,
                    Spec,
                    Grid,
                    DataStores,
                    std::enable_if_t<(tuple_util::size<DataStores>::value >= 3) &&
                                     tuple_util::size<DataStores>::value % 2 == 1>> {
                    using factor_t = std::integral_constant<size_t, (tuple_util::size<DataStores>::value - 1) / 2>;
                    struct comp_t {
                        template <class O, class I, class R>
                        auto operator()(O out, I in, R rho) const {
                            return execute_parallel().stage(F(), out, in, rho);
                        }
                    };
                    template <size_t I>
                    using x_arg = expandalble_frontend_impl_::arg<I>;
                    using fe_spec_t = decltype(std::declval<comp_t>()(expandalble_frontend_impl_::expandable<x_arg<0>>(),
                        expandalble_frontend_impl_::expandable<x_arg<1>>(),
                        x_arg<2>()));
                    using expected_t = core::convert_fe_to_be_spec<
                        expandalble_frontend_impl_::expand_spec<factor_t, fe_spec_t>,
                        typename Grid::interval_t,
                        DataStores>;
                    static constexpr bool value = std::is_same<normalized<Spec>, normalized<expected_t>>::value;
                };
            } // namespace shape
        } // namespace b200_backend
    } // namespace stencil
} // namespace gridtools
