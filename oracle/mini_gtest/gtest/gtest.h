// Minimal stand-in for <gtest/gtest.h> (GoogleTest is not in this image): just enough of TEST / TEST_F / EXPECT_* / ASSERT_* to compile the
// reference's own unit-test sources UNMODIFIED, once against the reference build (oracle/_ref) and once against the B200 host mirror
// (alphazero-multi-game_b200/host), and to print one line per test.  Test infrastructure only (oracle/run_ref_unit_tests.sh).
// Every test runs in a forked child, so a test that crashes the process (the reference's Renju code does) is reported as CRASH and the
// run goes on.  Output: "<suite>.<name> PASS|FAIL|CRASH", then "summary pass=.. fail=.. crash=..".
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <sys/wait.h>
#include <unistd.h>

namespace testing {
class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};
struct Registry {
    struct Entry { std::string name; std::function<Test*()> make; };
    static std::vector<Entry>& all() { static std::vector<Entry> v; return v; }
    static int& failures() { static int f = 0; return f; }
    static bool add(const char* suite, const char* name, std::function<Test*()> make) { all().push_back({std::string(suite) + "." + name, make}); return true; }
};
struct Message {            // swallows the `<< "text"` that may follow an assertion
    template <class T> Message& operator<<(const T&) { return *this; }
    Message& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
inline Message fail_at(const char* file, int line, const char* what) {
    ++Registry::failures();
    std::fprintf(stderr, "    %s:%d: %s\n", file, line, what);
    return Message();
}
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing

inline int RUN_ALL_TESTS() {
    int pass = 0, fail = 0, crash = 0;
    for (auto& e : testing::Registry::all()) {
        std::fflush(stdout); std::fflush(stderr);
        const pid_t pid = fork();
        if (pid == 0) {
            int rc = 0;
            try {
                testing::Test* t = e.make();
                t->SetUp(); t->TestBody(); t->TearDown();
                delete t;
                rc = testing::Registry::failures() ? 1 : 0;
            } catch (const std::exception& ex) { std::fprintf(stderr, "    uncaught exception: %s\n", ex.what()); rc = 1; }
            catch (...) { std::fprintf(stderr, "    uncaught exception\n"); rc = 1; }
            std::fflush(stdout); std::fflush(stderr);
            _exit(rc);
        }
        int st = 0; waitpid(pid, &st, 0);
        const char* verdict = (WIFEXITED(st) && WEXITSTATUS(st) == 0) ? "PASS" : WIFEXITED(st) ? "FAIL" : "CRASH";
        if (verdict[0] == 'P') ++pass; else if (verdict[0] == 'F') ++fail; else ++crash;
        std::printf("%s %s\n", e.name.c_str(), verdict);
    }
    std::printf("summary pass=%d fail=%d crash=%d\n", pass, fail, crash);
    return 0;
}

#define MG_CAT_(a, b) a##b
#define MG_CAT(a, b) MG_CAT_(a, b)
#define MG_TEST_(suite, name, base)                                                                              \
    class MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test))) : public base { public: void TestBody() override; };      \
    static bool MG_CAT(mg_reg_, MG_CAT(suite, MG_CAT(_, name))) = ::testing::Registry::add(#suite, #name,        \
        []() -> ::testing::Test* { return new MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test)))(); });                \
    void MG_CAT(suite, MG_CAT(_, MG_CAT(name, _Test)))::TestBody()
#define TEST(suite, name) MG_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) MG_TEST_(fixture, name, fixture)

#define MG_CHECK_(cond, text, fatal) \
    if (cond) ; else if (fatal) return (void)::testing::fail_at(__FILE__, __LINE__, text); else ::testing::fail_at(__FILE__, __LINE__, text)
#define EXPECT_TRUE(c) MG_CHECK_((c), "EXPECT_TRUE(" #c ")", false)
#define EXPECT_FALSE(c) MG_CHECK_(!(c), "EXPECT_FALSE(" #c ")", false)
#define ASSERT_TRUE(c) MG_CHECK_((c), "ASSERT_TRUE(" #c ")", true)
#define ASSERT_FALSE(c) MG_CHECK_(!(c), "ASSERT_FALSE(" #c ")", true)
#define EXPECT_EQ(a, b) MG_CHECK_(((a) == (b)), "EXPECT_EQ(" #a ", " #b ")", false)
#define EXPECT_NE(a, b) MG_CHECK_(((a) != (b)), "EXPECT_NE(" #a ", " #b ")", false)
#define EXPECT_GT(a, b) MG_CHECK_(((a) > (b)), "EXPECT_GT(" #a ", " #b ")", false)
#define EXPECT_GE(a, b) MG_CHECK_(((a) >= (b)), "EXPECT_GE(" #a ", " #b ")", false)
#define EXPECT_LT(a, b) MG_CHECK_(((a) < (b)), "EXPECT_LT(" #a ", " #b ")", false)
#define EXPECT_LE(a, b) MG_CHECK_(((a) <= (b)), "EXPECT_LE(" #a ", " #b ")", false)
#define ASSERT_EQ(a, b) MG_CHECK_(((a) == (b)), "ASSERT_EQ(" #a ", " #b ")", true)
#define ASSERT_NE(a, b) MG_CHECK_(((a) != (b)), "ASSERT_NE(" #a ", " #b ")", true)
#define EXPECT_FLOAT_EQ(a, b) MG_CHECK_((std::fabs((double)(a) - (double)(b)) <= 4e-7 * std::fmax(1.0, std::fmax(std::fabs((double)(a)), std::fabs((double)(b))))), "EXPECT_FLOAT_EQ(" #a ", " #b ")", false)
#define EXPECT_NEAR(a, b, tol) MG_CHECK_((std::fabs((double)(a) - (double)(b)) <= (double)(tol)), "EXPECT_NEAR(" #a ", " #b ")", false)
#define EXPECT_THROW(stmt, ex) \
    if ([&]() { try { stmt; } catch (const ex&) { return true; } catch (...) { return false; } return false; }()) ; else ::testing::fail_at(__FILE__, __LINE__, "EXPECT_THROW(" #stmt ", " #ex ")")
#define EXPECT_NO_THROW(stmt) \
    if ([&]() { try { stmt; } catch (...) { return false; } return true; }()) ; else ::testing::fail_at(__FILE__, __LINE__, "EXPECT_NO_THROW(" #stmt ")")
