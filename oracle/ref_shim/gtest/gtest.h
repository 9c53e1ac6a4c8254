// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for <gtest/gtest.h> (googletest is not in this
// image), just enough to run the REFERENCE'S OWN test files unmodified against the oracle build
// (the reference's sources over oracle/ref_shim): TEST / GTEST_TEST, value-parameterised suites
// (TestWithParam, TEST_P, INSTANTIATE_TEST_SUITE_P with testing::Values), the EXPECT_ / ASSERT_
// comparisons those files use, InitGoogleTest and RUN_ALL_TESTS. Failures are counted and
// printed; RUN_ALL_TESTS returns 1 when any expectation failed.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing
{
namespace internal
{
struct State
{
  int failures = 0;
  int expectations = 0;
  bool fatal = false;  // an ASSERT_ failed in the running test
  static State& Get()
  {
    static State state;
    return state;
  }
};

struct PlainTest
{
  std::string name;
  std::function<void()> body;
};
inline std::vector<PlainTest>& PlainTests()
{
  static std::vector<PlainTest> tests;
  return tests;
}

// per suite: the TEST_P bodies (as functions of a type-erased parameter index) and the
// instantiations (each runs every body once per value)
struct ParamSuite
{
  std::vector<std::pair<std::string, std::function<void(size_t, size_t)>>> bodies;
  std::vector<std::pair<std::string, size_t>> instantiations;  // name, number of values
};
inline std::map<std::string, ParamSuite>& ParamSuites()
{
  static std::map<std::string, ParamSuite> suites;
  return suites;
}

// 4-ULP comparison, as googletest's EXPECT_FLOAT_EQ / EXPECT_DOUBLE_EQ
template <typename Float, typename Bits>
inline bool AlmostEqual(Float a, Float b)
{
  if (std::isnan(a) || std::isnan(b))
  {
    return false;
  }
  if (a == b)
  {
    return true;
  }
  Bits ia, ib;
  std::memcpy(&ia, &a, sizeof(Float));
  std::memcpy(&ib, &b, sizeof(Float));
  constexpr Bits kSign = static_cast<Bits>(1) << (sizeof(Bits) * 8 - 1);
  const auto biased = [](Bits bits) { return (bits & kSign) ? ~bits + 1 : (kSign | bits); };
  const Bits ba = biased(ia), bb = biased(ib);
  return (ba >= bb ? ba - bb : bb - ba) <= 4;
}

template <typename A, typename B>
inline void Report(bool ok, const char* what, const A& a, const B& b, const char* file, int line,
                   bool fatal)
{
  State& state = State::Get();
  state.expectations++;
  if (!ok)
  {
    state.failures++;
    state.fatal = state.fatal || fatal;
    std::ostringstream text;
    text << file << ":" << line << ": " << what << " failed";
    std::cout << text.str() << std::endl;
    (void)a;
    (void)b;
  }
}
}  // namespace internal

class Test
{
public:
  virtual ~Test() {}
  virtual void TestBody() = 0;
};

template <typename T>
class TestWithParam : public Test
{
public:
  using ParamType = T;
  const T& GetParam() const { return *Parameter(); }
  static const T*& Parameter()
  {
    static const T* parameter = nullptr;
    return parameter;
  }
};

// testing::Values(...) -> a vector of the suite's parameter type, built at instantiation
template <typename... Ts>
struct ValueList
{
  std::tuple<Ts...> values;
};
template <typename... Ts>
inline ValueList<Ts...> Values(Ts... values)
{
  return ValueList<Ts...>{std::tuple<Ts...>(values...)};
}

namespace internal
{
template <typename T>
inline std::vector<std::vector<T>>& ParamValues(const std::string& suite)
{
  static std::map<std::string, std::vector<std::vector<T>>> values;
  return values[suite];
}

template <typename T, typename Tuple, size_t... I>
inline std::vector<T> ToVector(const Tuple& tuple, std::index_sequence<I...>)
{
  return std::vector<T>{T(std::get<I>(tuple))...};
}

template <typename Suite, typename... Ts>
inline int Instantiate(const char* prefix, const char* suite, const ValueList<Ts...>& list)
{
  using T = typename Suite::ParamType;
  auto& per_suite = ParamValues<T>(suite);
  per_suite.push_back(ToVector<T>(list.values, std::index_sequence_for<Ts...>{}));
  ParamSuites()[suite].instantiations.emplace_back(prefix, sizeof...(Ts));
  return 0;
}

template <typename Fixture>
inline int RegisterParamBody(const char* suite, const char* name)
{
  using T = typename Fixture::ParamType;
  ParamSuites()[suite].bodies.emplace_back(
      name, [suite](size_t instantiation, size_t value)
      {
        const T& parameter = ParamValues<T>(suite)[instantiation][value];
        TestWithParam<T>::Parameter() = &parameter;
        Fixture fixture;
        fixture.TestBody();
      });
  return 0;
}

inline int RegisterPlain(const char* suite, const char* name, std::function<void()> body)
{
  PlainTests().push_back(PlainTest{std::string(suite) + "." + name, std::move(body)});
  return 0;
}
}  // namespace internal

inline void InitGoogleTest(int*, char**) {}

inline int RunAllTests()
{
  using namespace internal;
  int ran = 0;
  for (const PlainTest& test : PlainTests())
  {
    State::Get().fatal = false;
    std::cout << "[ RUN ] " << test.name << std::endl;
    test.body();
    ran++;
  }
  for (auto& suite : ParamSuites())
  {
    for (size_t i = 0; i < suite.second.instantiations.size(); i++)
    {
      for (size_t v = 0; v < suite.second.instantiations[i].second; v++)
      {
        for (auto& body : suite.second.bodies)
        {
          State::Get().fatal = false;
          std::cout << "[ RUN ] " << suite.second.instantiations[i].first << "/" << suite.first
                    << "." << body.first << "/" << v << std::endl;
          body.second(i, v);
          ran++;
        }
      }
    }
  }
  const State& state = State::Get();
  std::cout << "[ DONE ] " << ran << " tests, " << state.expectations << " expectations, "
            << state.failures << " failed" << std::endl;
  return (state.failures == 0 && ran > 0) ? 0 : 1;
}
}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define VGT_GTEST_CLASS_(suite, name) suite##_##name##_Test
#define VGT_GTEST_BODY_(suite, name) suite##_##name##_Test_body
#define VGT_GTEST_FLAG_(suite, name) suite##_##name##_Test_registered

#define GTEST_TEST(suite, name)                                                         \
  static void VGT_GTEST_BODY_(suite, name)();                                           \
  static const int VGT_GTEST_FLAG_(suite, name) =                                       \
      ::testing::internal::RegisterPlain(#suite, #name, &VGT_GTEST_BODY_(suite, name)); \
  static void VGT_GTEST_BODY_(suite, name)()
#define TEST(suite, name) GTEST_TEST(suite, name)

#define TEST_P(suite, name)                                                                 \
  class VGT_GTEST_CLASS_(suite, name) : public suite                                        \
  {                                                                                         \
  public:                                                                                   \
    void TestBody() override;                                                               \
  };                                                                                        \
  static const int VGT_GTEST_FLAG_(suite, name) =                                           \
      ::testing::internal::RegisterParamBody<VGT_GTEST_CLASS_(suite, name)>(#suite, #name); \
  void VGT_GTEST_CLASS_(suite, name)::TestBody()

#define INSTANTIATE_TEST_SUITE_P(prefix, suite, values)             \
  static const int prefix##_##suite##_instantiated =                \
      ::testing::internal::Instantiate<suite>(#prefix, #suite, values)

#define VGT_GTEST_COMPARE_(a, b, op, text, fatal)                                            \
  do                                                                                         \
  {                                                                                          \
    const auto& vgt_a__ = (a);                                                               \
    const auto& vgt_b__ = (b);                                                               \
    ::testing::internal::Report((vgt_a__ op vgt_b__), text, vgt_a__, vgt_b__, __FILE__,      \
                                __LINE__, fatal);                                            \
  } while (0)

#define EXPECT_EQ(a, b) VGT_GTEST_COMPARE_(a, b, ==, "EXPECT_EQ(" #a ", " #b ")", false)
#define EXPECT_NE(a, b) VGT_GTEST_COMPARE_(a, b, !=, "EXPECT_NE(" #a ", " #b ")", false)
#define EXPECT_LT(a, b) VGT_GTEST_COMPARE_(a, b, <, "EXPECT_LT(" #a ", " #b ")", false)
#define EXPECT_LE(a, b) VGT_GTEST_COMPARE_(a, b, <=, "EXPECT_LE(" #a ", " #b ")", false)
#define EXPECT_GT(a, b) VGT_GTEST_COMPARE_(a, b, >, "EXPECT_GT(" #a ", " #b ")", false)
#define EXPECT_GE(a, b) VGT_GTEST_COMPARE_(a, b, >=, "EXPECT_GE(" #a ", " #b ")", false)
#define EXPECT_TRUE(c)                                                                       \
  ::testing::internal::Report(static_cast<bool>(c), "EXPECT_TRUE(" #c ")", 0, 0, __FILE__,   \
                              __LINE__, false)
#define EXPECT_FALSE(c)                                                                      \
  ::testing::internal::Report(!static_cast<bool>(c), "EXPECT_FALSE(" #c ")", 0, 0, __FILE__, \
                              __LINE__, false)
#define EXPECT_FLOAT_EQ(a, b)                                                                \
  ::testing::internal::Report(                                                               \
      ::testing::internal::AlmostEqual<float, uint32_t>(static_cast<float>(a),               \
                                                        static_cast<float>(b)),              \
      "EXPECT_FLOAT_EQ(" #a ", " #b ")", 0, 0, __FILE__, __LINE__, false)
#define EXPECT_DOUBLE_EQ(a, b)                                                               \
  ::testing::internal::Report(                                                               \
      ::testing::internal::AlmostEqual<double, uint64_t>(static_cast<double>(a),             \
                                                         static_cast<double>(b)),            \
      "EXPECT_DOUBLE_EQ(" #a ", " #b ")", 0, 0, __FILE__, __LINE__, false)
// (an ASSERT_ that fails ends the test body, like googletest's)
#define ASSERT_EQ(a, b)                                                                      \
  do                                                                                         \
  {                                                                                          \
    VGT_GTEST_COMPARE_(a, b, ==, "ASSERT_EQ(" #a ", " #b ")", true);                         \
    if (::testing::internal::State::Get().fatal) { return; }                                 \
  } while (0)
#define ASSERT_TRUE(c)                                                                       \
  do                                                                                         \
  {                                                                                          \
    ::testing::internal::Report(static_cast<bool>(c), "ASSERT_TRUE(" #c ")", 0, 0, __FILE__, \
                                __LINE__, true);                                             \
    if (::testing::internal::State::Get().fatal) { return; }                                 \
  } while (0)
#define ASSERT_NO_THROW(statement)                                                           \
  do                                                                                         \
  {                                                                                          \
    bool vgt_threw__ = false;                                                                \
    try { statement; } catch (...) { vgt_threw__ = true; }                                   \
    ::testing::internal::Report(!vgt_threw__, "ASSERT_NO_THROW(" #statement ")", 0, 0,       \
                                __FILE__, __LINE__, true);                                   \
    if (::testing::internal::State::Get().fatal) { return; }                                 \
  } while (0)
