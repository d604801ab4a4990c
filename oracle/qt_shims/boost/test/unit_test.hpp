// Minimal stand-in for Boost.Test (not installed in this image): just what the reference's
// test/test_gpusim.cpp uses — BOOST_AUTO_TEST_CASE, BOOST_CHECK_EQUAL, BOOST_REQUIRE_EQUAL — and a
// main() that runs the registered cases (optionally only those named by --run_test=A,B).
// TEST INFRASTRUCTURE for `make refcheck`.  C strings compare by content, as in Boost.Test.
#pragma once
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace gsb_boost_stub
{
struct Case {
    const char* name;
    void (*fn)();
};
inline std::vector<Case>& cases()
{
    static std::vector<Case> v;
    return v;
}
inline int& failures()
{
    static int n = 0;
    return n;
}
struct Registrar {
    Registrar(const char* name, void (*fn)()) { cases().push_back({name, fn}); }
};
struct RequireFailed : std::exception {
};
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wsign-compare" // the reference compares size() with int counts
template <class A, class B> bool equal(const A& a, const B& b) { return a == b; }
#pragma GCC diagnostic pop
inline bool equal(const char* a, const char* b) { return (a && b) ? std::strcmp(a, b) == 0 : a == b; }
inline bool equal(char* a, char* b) { return equal(static_cast<const char*>(a), static_cast<const char*>(b)); }
inline bool equal(char* a, const char* b) { return equal(static_cast<const char*>(a), b); }
inline bool equal(const char* a, char* b) { return equal(a, static_cast<const char*>(b)); }
template <class A, class B>
void check(const A& a, const B& b, const char* ea, const char* eb, const char* file, int line, bool require)
{
    if (equal(a, b))
        return;
    failures()++;
    std::cerr << file << "(" << line << "): error: check " << ea << " == " << eb << " has failed [" << a << " != " << b
              << "]" << std::endl;
    if (require)
        throw RequireFailed();
}
} // namespace gsb_boost_stub

#define BOOST_AUTO_TEST_CASE(name)                                                               \
    static void name##_body();                                                                   \
    static gsb_boost_stub::Registrar name##_registrar(#name, &name##_body);                      \
    static void name##_body()
#define BOOST_CHECK_EQUAL(a, b) gsb_boost_stub::check((a), (b), #a, #b, __FILE__, __LINE__, false)
#define BOOST_REQUIRE_EQUAL(a, b) gsb_boost_stub::check((a), (b), #a, #b, __FILE__, __LINE__, true)

#ifdef BOOST_TEST_MODULE
int main(int argc, char** argv)
{
    std::string only;
    for (int i = 1; i < argc; i++)
        if (std::strncmp(argv[i], "--run_test=", 11) == 0)
            only = std::string(",") + (argv[i] + 11) + ",";
    int ran = 0;
    for (const auto& c : gsb_boost_stub::cases()) {
        if (!only.empty() && only.find(std::string(",") + c.name + ",") == std::string::npos)
            continue;
        std::cout << "Entering test case \"" << c.name << "\"" << std::endl;
        const int before = gsb_boost_stub::failures();
        try {
            c.fn();
        } catch (const gsb_boost_stub::RequireFailed&) {
        } catch (const std::exception& e) {
            gsb_boost_stub::failures()++;
            std::cerr << "error: exception in \"" << c.name << "\": " << e.what() << std::endl;
        }
        std::cout << "Leaving test case \"" << c.name << "\"" << (gsb_boost_stub::failures() == before ? "" : " (FAILED)")
                  << std::endl;
        ran++;
    }
    if (gsb_boost_stub::failures() == 0)
        std::cout << "\n*** No errors detected (" << ran << " test cases)" << std::endl;
    else
        std::cout << "\n*** " << gsb_boost_stub::failures() << " failure(s) detected in " << ran << " test cases" << std::endl;
    return gsb_boost_stub::failures() == 0 ? 0 : 201;
}
#endif
