/* gmock/gmock.h -- the two matchers the reference's gcl test uses (ASSERT_THAT / EXPECT_THAT with ContainerEq, Eq).
 * TEST INFRASTRUCTURE ONLY, see ../gtest/gtest.h. */
#ifndef GTB200_GMOCK_SHIM_H
#define GTB200_GMOCK_SHIM_H

#include <iterator>

#include "../gtest/gtest.h"

namespace testing {
    namespace internal {
        template <class C>
        struct container_eq_matcher {
            C const &expected;
            template <class A>
            AssertionResult match(A const &actual) const {
                auto a = std::begin(actual), ae = std::end(actual);
                auto e = std::begin(expected), ee = std::end(expected);
                size_t i = 0;
                for (; a != ae && e != ee; ++a, ++e, ++i)
                    if (!elem_eq(*a, *e))
                        return AssertionResult(false, "containers differ at element " + std::to_string(i));
                if (a != ae || e != ee)
                    return AssertionResult(false, "containers have different sizes");
                return AssertionSuccess();
            }
            template <class X, class Y>
            static auto elem_eq(X const &x, Y const &y) -> decltype(bool(x == y)) {
                return x == y;
            }
            template <class X, class Y, size_t N>
            static bool elem_eq(X const (&x)[N], Y const (&y)[N]) { // rows of a C array (halos[f] in the gcl test)
                for (size_t i = 0; i < N; ++i)
                    if (!elem_eq(x[i], y[i]))
                        return false;
                return true;
            }
        };
        template <class V>
        struct eq_matcher {
            V expected;
            template <class A>
            AssertionResult match(A const &actual) const {
                return actual == expected ? AssertionSuccess()
                                          : AssertionResult(false, "value differs: " + print_value(actual));
            }
        };
    } // namespace internal
    template <class C>
    internal::container_eq_matcher<C> ContainerEq(C const &c) {
        return {c};
    }
    template <class V>
    internal::eq_matcher<V> Eq(V v) {
        return {v};
    }
} // namespace testing

#define EXPECT_THAT(value, matcher) GTEST_SHIM_CHECK_((matcher).match(value), GTEST_SHIM_NONFATAL_)
#define ASSERT_THAT(value, matcher) GTEST_SHIM_CHECK_((matcher).match(value), GTEST_SHIM_FATAL_)

#endif
