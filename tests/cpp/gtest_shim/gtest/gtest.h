/*
 * gtest/gtest.h -- a small stand-in for the googletest interface the reference's regression tests use.
 *
 * TEST INFRASTRUCTURE ONLY.  googletest is fetched from GitHub by the reference's build (cmake/internal/
 * FetchGoogletest.cmake) and is not installed in this image, so the reference's own test sources
 * (tests/regression/*.cpp, tests/regression/gcl/test_halo_exchange_3D.cpp, tests/src/regression_main.cpp) could not
 * be compiled unchanged.  This header implements the subset they need: TEST / TEST_F / TEST_P /
 * INSTANTIATE_TEST_SUITE_P / TYPED_TEST_SUITE / TYPED_TEST, EXPECT_* / ASSERT_* with message streaming,
 * testing::Values / Types / Test / TestWithParam, --gtest_filter patterns, InitGoogleTest / RUN_ALL_TESTS.
 * Thread safe as far as the in-process MPI stand-in needs it (every "rank" thread may run RUN_ALL_TESTS()).
 */
#ifndef GTB200_GTEST_SHIM_H
#define GTB200_GTEST_SHIM_H

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <tuple>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <vector>

namespace testing {
    inline std::string FLAGS_gtest_filter = "*";

    class Test {
      public:
        virtual ~Test() {}
        virtual void SetUp() {}
        virtual void TearDown() {}
        virtual void TestBody() = 0;
    };

    template <class T>
    class WithParamInterface {
      public:
        using ParamType = T;
        static const T &GetParam() { return *param_ptr(); }
        static const T *&param_ptr() {
            static const T *p = nullptr; // shared by the rank threads of the MPI stand-in: see state_t::between_tests
            return p;
        }
    };

    template <class T>
    class TestWithParam : public Test, public WithParamInterface<T> {};

    template <class... Ts>
    struct Types {};

    class Message {
        std::ostringstream m_ss;

      public:
        Message() {}
        Message(Message const &o) { m_ss << o.str(); }
        template <class T>
        Message &operator<<(T const &v) {
            m_ss << v;
            return *this;
        }
        Message &operator<<(std::ostream &(*f)(std::ostream &)) {
            m_ss << f;
            return *this;
        }
        std::string str() const { return m_ss.str(); }
    };

    class AssertionResult {
        bool m_ok;
        std::string m_msg;

      public:
        AssertionResult(bool ok, std::string msg = {}) : m_ok(ok), m_msg(std::move(msg)) {}
        explicit operator bool() const { return m_ok; }
        std::string const &message() const { return m_msg; }
        template <class T>
        AssertionResult &operator<<(T const &v) {
            std::ostringstream ss;
            ss << v;
            m_msg += ss.str();
            return *this;
        }
    };
    inline AssertionResult AssertionSuccess() { return AssertionResult(true); }
    inline AssertionResult AssertionFailure() { return AssertionResult(false); }

    namespace internal {
        struct test_info {
            std::string suite, name;
            std::function<Test *()> factory;
            std::function<void()> before; // sets the parameter of a TEST_P instance
        };
        struct state_t {
            std::vector<test_info> tests;
            std::vector<std::function<void()>> finalizers; // expand TEST_P x INSTANTIATE at RUN_ALL_TESTS
            std::atomic<int> failures{0};
            std::mutex print;
            bool quiet = false;
            // called before every test; the MPI stand-in puts a barrier here, so that all rank threads are in the
            // same test (the parameter pointer of TEST_P is one static per suite, like in googletest)
            std::function<void()> between_tests;
        };
        inline state_t &state() {
            static state_t s;
            return s;
        }
        inline int &current_failures() { // per thread: failures of the running test
            static thread_local int n = 0;
            return n;
        }

        template <class T, class = void>
        struct streamable : std::false_type {};
        template <class T>
        struct streamable<T, std::void_t<decltype(std::declval<std::ostream &>() << std::declval<T const &>())>>
            : std::true_type {};

        template <class T>
        std::string print_value(T const &v) {
            std::ostringstream ss;
            if constexpr (std::is_same_v<T, bool>)
                ss << (v ? "true" : "false");
            else if constexpr (streamable<T>::value) {
                if constexpr (std::is_floating_point_v<T>)
                    ss.precision(17);
                ss << v;
            } else {
                ss << sizeof(T) << "-byte object <";
                auto *p = reinterpret_cast<unsigned char const *>(&v);
                for (size_t i = 0; i < sizeof(T) && i < 32; ++i) {
                    char b[4];
                    std::snprintf(b, sizeof b, "%02X ", p[i]);
                    ss << b;
                }
                ss << ">";
            }
            return ss.str();
        }

        class AssertHelper {
            const char *m_file;
            int m_line;
            std::string m_msg;

          public:
            AssertHelper(const char *file, int line, std::string msg) : m_file(file), m_line(line), m_msg(std::move(msg)) {}
            void operator=(Message const &user) const {
                state_t &s = state();
                ++s.failures;
                ++current_failures();
                std::lock_guard<std::mutex> l(s.print);
                std::cout << m_file << ":" << m_line << ": Failure\n" << m_msg;
                std::string u = user.str();
                if (!u.empty())
                    std::cout << "\n" << u;
                std::cout << std::endl;
            }
        };

        template <class A, class B, class Op>
        AssertionResult compare(const char *ea, const char *eb, A const &a, B const &b, Op op, const char *opname) {
            if (op(a, b))
                return AssertionSuccess();
            return AssertionResult(false,
                std::string("Expected: (") + ea + ") " + opname + " (" + eb + "), actual: " + print_value(a) + " vs " +
                    print_value(b));
        }
        struct eq_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return a == b;
            }
        };
        struct ne_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return !(a == b);
            }
        };
        struct lt_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return a < b;
            }
        };
        struct le_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return a <= b;
            }
        };
        struct gt_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return a > b;
            }
        };
        struct ge_op {
            template <class A, class B>
            bool operator()(A const &a, B const &b) const {
                return a >= b;
            }
        };
        inline AssertionResult near(const char *ea, const char *eb, double a, double b, double tol) {
            if (std::fabs(a - b) <= tol)
                return AssertionSuccess();
            return AssertionResult(false,
                std::string("The difference between ") + ea + " and " + eb + " is " + print_value(std::fabs(a - b)) +
                    ", which exceeds " + print_value(tol) + " (" + print_value(a) + " vs " + print_value(b) + ")");
        }
        template <class T>
        AssertionResult almost_equal(const char *ea, const char *eb, T a, T b) { // 4 ULPs like googletest
            if (a == b)
                return AssertionSuccess();
            T d = std::fabs(a - b), m = std::fmax(std::fabs(a), std::fabs(b));
            if (d <= 4 * std::numeric_limits<T>::epsilon() * m)
                return AssertionSuccess();
            return AssertionResult(false,
                std::string("Expected equality of ") + ea + " and " + eb + ": " + print_value(a) + " vs " + print_value(b));
        }
        inline AssertionResult boolean(const char *e, bool want, AssertionResult const &r) {
            if (bool(r) == want)
                return AssertionSuccess();
            return AssertionResult(false,
                std::string("Value of: ") + e + "\n  Actual: " + (want ? "false" : "true") +
                    (r.message().empty() ? "" : " (" + r.message() + ")") + "\nExpected: " + (want ? "true" : "false"));
        }
        inline AssertionResult boolean(const char *e, bool want, bool v) { return boolean(e, want, AssertionResult(v)); }
        template <class T, std::enable_if_t<!std::is_same_v<T, AssertionResult> && !std::is_same_v<T, bool>, int> = 0>
        AssertionResult boolean(const char *e, bool want, T const &v) {
            return boolean(e, want, AssertionResult(static_cast<bool>(v)));
        }

        inline int add_test(std::string suite, std::string name, std::function<Test *()> f, std::function<void()> before = {}) {
            state().tests.push_back({std::move(suite), std::move(name), std::move(f), std::move(before)});
            return 0;
        }

        // ---- typed tests
        struct default_name_generator {
            template <class T>
            static std::string GetName(int i) {
                return std::to_string(i);
            }
        };
        template <class G = default_name_generator>
        struct name_generator_selector {
            using type = G;
        };
        template <template <class> class Fixture, class NameGen, class... Ts>
        int register_typed(const char *suite, const char *name, Types<Ts...>) {
            int i = 0;
            (void)std::initializer_list<int>{(add_test(std::string(suite) + "/" + NameGen::template GetName<Ts>(i), name,
                                                  [] { return static_cast<Test *>(new Fixture<Ts>); }),
                ++i)...};
            return 0;
        }

        // ---- value-parameterised tests
        template <class Suite>
        struct param_registry {
            using param_t = typename Suite::ParamType;
            struct inst {
                std::string prefix;
                std::vector<param_t> values;
            };
            static std::vector<std::pair<std::string, std::function<Test *()>>> &tests() {
                static std::vector<std::pair<std::string, std::function<Test *()>>> v;
                return v;
            }
            static std::vector<inst> &insts() {
                static std::vector<inst> v;
                return v;
            }
            static void hook(const char *suite) {
                static bool done = false;
                if (done)
                    return;
                done = true;
                std::string s = suite;
                state().finalizers.push_back([s] {
                    for (auto &in : insts())
                        for (auto &t : tests())
                            for (size_t i = 0; i < in.values.size(); ++i) {
                                const param_t *p = &in.values[i];
                                add_test(in.prefix + "/" + s, t.first + "/" + std::to_string(i), t.second,
                                    [p] { Suite::param_ptr() = p; });
                            }
                });
            }
            static int add(const char *suite, const char *name, std::function<Test *()> f) {
                hook(suite);
                tests().emplace_back(name, std::move(f));
                return 0;
            }
            template <class Gen>
            static int instantiate(const char *suite, const char *prefix, Gen const &g) {
                hook(suite);
                insts().push_back({prefix, g.template as<param_t>()});
                return 0;
            }
        };
        template <class... Ts>
        struct value_array {
            std::tuple<Ts...> v;
            template <class P>
            std::vector<P> as() const {
                return std::apply([](auto const &...x) { return std::vector<P>{static_cast<P>(x)...}; }, v);
            }
        };

        inline bool glob(const char *p, const char *s) {
            if (!*p)
                return !*s;
            if (*p == '*')
                return glob(p + 1, s) || (*s && glob(p, s + 1));
            return *s && (*p == '?' || *p == *s) && glob(p + 1, s + 1);
        }
        inline bool matches_any(std::string const &pats, std::string const &name) {
            size_t pos = 0;
            while (pos <= pats.size()) {
                size_t c = pats.find(':', pos);
                std::string one = pats.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
                if (!one.empty() && glob(one.c_str(), name.c_str()))
                    return true;
                if (c == std::string::npos)
                    break;
                pos = c + 1;
            }
            return false;
        }
        inline bool selected(std::string const &name) {
            std::string f = FLAGS_gtest_filter, pos = f, neg;
            size_t dash = f.find('-');
            if (dash != std::string::npos) {
                pos = f.substr(0, dash);
                neg = f.substr(dash + 1);
            }
            if (pos.empty())
                pos = "*";
            return matches_any(pos, name) && !matches_any(neg, name);
        }
    } // namespace internal

    template <class... Ts>
    internal::value_array<Ts...> Values(Ts... v) {
        return {std::make_tuple(v...)};
    }

    // the little of the listener API tests/src/regression_main.cpp touches (dropping the default printer = quiet)
    class TestEventListener {
      public:
        virtual ~TestEventListener() {}
    };
    class EmptyTestEventListener : public TestEventListener {};
    class TestEventListeners {
        TestEventListener m_default;

      public:
        TestEventListener *default_result_printer() { return &m_default; }
        TestEventListener *Release(TestEventListener *) {
            internal::state().quiet = true;
            return nullptr;
        }
        void Append(TestEventListener *) {}
    };
    class UnitTest {
        TestEventListeners m_listeners;

      public:
        static UnitTest *GetInstance() {
            static UnitTest u;
            return &u;
        }
        TestEventListeners &listeners() { return m_listeners; }
    };

    inline void InitGoogleTest(int *argc, char **argv) {
        int out = 1;
        for (int i = 1; i < *argc; ++i) {
            if (std::strncmp(argv[i], "--gtest_filter=", 15) == 0)
                FLAGS_gtest_filter = argv[i] + 15;
            else if (std::strncmp(argv[i], "--gtest_", 8) == 0) {
            } else
                argv[out++] = argv[i];
        }
        *argc = out;
    }
    inline void InitGoogleTest() {}

    inline int run_all_tests() {
        auto &s = internal::state();
        static std::once_flag once;
        std::call_once(once, [&] {
            for (auto &f : s.finalizers)
                f();
        });
        int ran = 0, failed = 0;
        std::vector<std::string> failed_names;
        for (auto &t : s.tests) {
            std::string full = t.suite + "." + t.name;
            if (!internal::selected(full))
                continue;
            if (!s.quiet) {
                std::lock_guard<std::mutex> l(s.print);
                std::cout << "[ RUN      ] " << full << std::endl;
            }
            internal::current_failures() = 0;
            if (s.between_tests)
                s.between_tests();
            if (t.before)
                t.before();
            {
                std::unique_ptr<Test> obj(t.factory());
                obj->SetUp();
                if (internal::current_failures() == 0)
                    obj->TestBody();
                obj->TearDown();
            }
            ++ran;
            bool ok = internal::current_failures() == 0;
            if (!ok) {
                ++failed;
                failed_names.push_back(full);
            }
            if (!s.quiet || !ok) {
                std::lock_guard<std::mutex> l(s.print);
                std::cout << (ok ? "[       OK ] " : "[  FAILED  ] ") << full << std::endl;
            }
        }
        std::lock_guard<std::mutex> l(s.print);
        std::cout << "[==========] " << ran << " tests ran.\n[  PASSED  ] " << ran - failed << " tests." << std::endl;
        if (failed) {
            std::cout << "[  FAILED  ] " << failed << " tests, listed below:\n";
            for (auto &n : failed_names)
                std::cout << "[  FAILED  ] " << n << "\n";
        }
        return failed ? 1 : 0;
    }
} // namespace testing

#define RUN_ALL_TESTS() ::testing::run_all_tests()

#define GTEST_SHIM_BLOCKER_ \
    switch (0)              \
    case 0:                 \
    default:
#define GTEST_SHIM_CHECK_(expr, on_fail) \
    GTEST_SHIM_BLOCKER_                  \
    if (const ::testing::AssertionResult gtest_ar_ = (expr)) \
        ;                                \
    else                                 \
        on_fail ::testing::internal::AssertHelper(__FILE__, __LINE__, gtest_ar_.message()) = ::testing::Message()
#define GTEST_SHIM_NONFATAL_
#define GTEST_SHIM_FATAL_ return

#define GTEST_SHIM_CMP_(a, b, op, name, on_fail) \
    GTEST_SHIM_CHECK_(::testing::internal::compare(#a, #b, a, b, ::testing::internal::op(), name), on_fail)

#define EXPECT_EQ(a, b) GTEST_SHIM_CMP_(a, b, eq_op, "==", GTEST_SHIM_NONFATAL_)
#define EXPECT_NE(a, b) GTEST_SHIM_CMP_(a, b, ne_op, "!=", GTEST_SHIM_NONFATAL_)
#define EXPECT_LT(a, b) GTEST_SHIM_CMP_(a, b, lt_op, "<", GTEST_SHIM_NONFATAL_)
#define EXPECT_LE(a, b) GTEST_SHIM_CMP_(a, b, le_op, "<=", GTEST_SHIM_NONFATAL_)
#define EXPECT_GT(a, b) GTEST_SHIM_CMP_(a, b, gt_op, ">", GTEST_SHIM_NONFATAL_)
#define EXPECT_GE(a, b) GTEST_SHIM_CMP_(a, b, ge_op, ">=", GTEST_SHIM_NONFATAL_)
#define ASSERT_EQ(a, b) GTEST_SHIM_CMP_(a, b, eq_op, "==", GTEST_SHIM_FATAL_)
#define ASSERT_NE(a, b) GTEST_SHIM_CMP_(a, b, ne_op, "!=", GTEST_SHIM_FATAL_)
#define ASSERT_LT(a, b) GTEST_SHIM_CMP_(a, b, lt_op, "<", GTEST_SHIM_FATAL_)
#define ASSERT_LE(a, b) GTEST_SHIM_CMP_(a, b, le_op, "<=", GTEST_SHIM_FATAL_)
#define ASSERT_GT(a, b) GTEST_SHIM_CMP_(a, b, gt_op, ">", GTEST_SHIM_FATAL_)
#define ASSERT_GE(a, b) GTEST_SHIM_CMP_(a, b, ge_op, ">=", GTEST_SHIM_FATAL_)
#define EXPECT_TRUE(c) GTEST_SHIM_CHECK_(::testing::internal::boolean(#c, true, c), GTEST_SHIM_NONFATAL_)
#define EXPECT_FALSE(c) GTEST_SHIM_CHECK_(::testing::internal::boolean(#c, false, c), GTEST_SHIM_NONFATAL_)
#define ASSERT_TRUE(c) GTEST_SHIM_CHECK_(::testing::internal::boolean(#c, true, c), GTEST_SHIM_FATAL_)
#define ASSERT_FALSE(c) GTEST_SHIM_CHECK_(::testing::internal::boolean(#c, false, c), GTEST_SHIM_FATAL_)
#define EXPECT_NEAR(a, b, t) GTEST_SHIM_CHECK_(::testing::internal::near(#a, #b, a, b, t), GTEST_SHIM_NONFATAL_)
#define ASSERT_NEAR(a, b, t) GTEST_SHIM_CHECK_(::testing::internal::near(#a, #b, a, b, t), GTEST_SHIM_FATAL_)
#define EXPECT_DOUBLE_EQ(a, b) \
    GTEST_SHIM_CHECK_(::testing::internal::almost_equal<double>(#a, #b, a, b), GTEST_SHIM_NONFATAL_)
#define EXPECT_FLOAT_EQ(a, b) \
    GTEST_SHIM_CHECK_(::testing::internal::almost_equal<float>(#a, #b, a, b), GTEST_SHIM_NONFATAL_)
#define ADD_FAILURE() GTEST_SHIM_CHECK_(::testing::AssertionResult(false, "Failed"), GTEST_SHIM_NONFATAL_)
#define FAIL() GTEST_SHIM_CHECK_(::testing::AssertionResult(false, "Failed"), GTEST_SHIM_FATAL_)
#define SUCCEED() GTEST_SHIM_CHECK_(::testing::AssertionResult(true), GTEST_SHIM_NONFATAL_)
#define EXPECT_THROW(stmt, ex)                                                                              \
    GTEST_SHIM_CHECK_(([&]() -> ::testing::AssertionResult {                                                \
        try {                                                                                               \
            stmt;                                                                                           \
        } catch (ex const &) {                                                                              \
            return ::testing::AssertionSuccess();                                                           \
        } catch (...) {                                                                                     \
            return ::testing::AssertionResult(false, #stmt " throws an exception of a different type");     \
        }                                                                                                   \
        return ::testing::AssertionResult(false, #stmt " throws nothing, expected " #ex);                   \
    })(),                                                                                                   \
        GTEST_SHIM_NONFATAL_)
#define EXPECT_NO_THROW(stmt)                                                        \
    GTEST_SHIM_CHECK_(([&]() -> ::testing::AssertionResult {                         \
        try {                                                                        \
            stmt;                                                                    \
        } catch (...) {                                                              \
            return ::testing::AssertionResult(false, #stmt " throws an exception");  \
        }                                                                            \
        return ::testing::AssertionSuccess();                                        \
    })(),                                                                            \
        GTEST_SHIM_NONFATAL_)

#define GTEST_SHIM_CLASS_(suite, name) suite##_##name##_Test

#define GTEST_SHIM_TEST_(suite, name, parent)                                                                        \
    class GTEST_SHIM_CLASS_(suite, name) : public parent {                                                           \
        void TestBody() override;                                                                                    \
        static int gtest_reg_;                                                                                       \
    };                                                                                                               \
    int GTEST_SHIM_CLASS_(suite, name)::gtest_reg_ = ::testing::internal::add_test(                                  \
        #suite, #name, [] { return static_cast<::testing::Test *>(new GTEST_SHIM_CLASS_(suite, name)); });           \
    void GTEST_SHIM_CLASS_(suite, name)::TestBody()

#define TEST(suite, name) GTEST_SHIM_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_SHIM_TEST_(fixture, name, fixture)

#define TEST_P(suite, name)                                                                                          \
    class GTEST_SHIM_CLASS_(suite, name) : public suite {                                                            \
        void TestBody() override;                                                                                    \
        static int gtest_reg_;                                                                                       \
    };                                                                                                               \
    int GTEST_SHIM_CLASS_(suite, name)::gtest_reg_ = ::testing::internal::param_registry<suite>::add(                \
        #suite, #name, [] { return static_cast<::testing::Test *>(new GTEST_SHIM_CLASS_(suite, name)); });           \
    void GTEST_SHIM_CLASS_(suite, name)::TestBody()

#define INSTANTIATE_TEST_SUITE_P(prefix, suite, ...)                 \
    static int gtest_##prefix##suite##_inst_ [[maybe_unused]] =      \
        ::testing::internal::param_registry<suite>::instantiate(#suite, #prefix, __VA_ARGS__)
#define INSTANTIATE_TEST_CASE_P INSTANTIATE_TEST_SUITE_P

#define TYPED_TEST_SUITE(suite, types, ...)            \
    typedef types gtest_type_params_##suite##_;        \
    typedef ::testing::internal::name_generator_selector<__VA_ARGS__>::type gtest_name_gen_##suite##_
#define TYPED_TEST_CASE TYPED_TEST_SUITE

#define TYPED_TEST(suite, name)                                                                                      \
    template <class gtest_TypeParam_>                                                                                \
    class GTEST_SHIM_CLASS_(suite, name) : public suite<gtest_TypeParam_> {                                          \
        typedef suite<gtest_TypeParam_> TestFixture;                                                                 \
        typedef gtest_TypeParam_ TypeParam;                                                                          \
        void TestBody() override;                                                                                    \
    };                                                                                                               \
    static int gtest_##suite##_##name##_reg_ [[maybe_unused]] =                                                      \
        ::testing::internal::register_typed<GTEST_SHIM_CLASS_(suite, name), gtest_name_gen_##suite##_>(              \
            #suite, #name, gtest_type_params_##suite##_());                                                          \
    template <class gtest_TypeParam_>                                                                                \
    void GTEST_SHIM_CLASS_(suite, name)<gtest_TypeParam_>::TestBody()

#endif
