// Header-only stand-ins for the handful of Qt types the reference's hot-path sources touch.
// TEST INFRASTRUCTURE: lets oracle/Makefile compile /root/reference/{calculation_functors.cpp,
// fingerprintdb_cuda.cpp,fingerprintdb_cuda.cu} verbatim (Qt5 is not installed in this image)
// so the reference itself can serve as parity checker and CPU/CUDA baseline.  Not Qt, not
// part of the product; only the members those three files use exist.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <future>
#include <iostream>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#ifdef __CUDACC__
// the reference .cu relies on transitive includes that CCCL 2.8 no longer provides
#include <thrust/functional.h>
#include <thrust/remove.h>
#endif

class QString
{
  public:
    QString() = default;
    QString(const char* s) : m_s(s ? s : "") {}
    QString(const std::string& s) : m_s(s) {}
    bool operator==(const QString& o) const { return m_s == o.m_s; }
    bool operator!=(const QString& o) const { return m_s != o.m_s; }
    bool operator<(const QString& o) const { return m_s < o.m_s; }
    const std::string& toStdString() const { return m_s; }

  private:
    std::string m_s;
};

class QObject
{
  public:
    virtual ~QObject() = default;
};

// qDebug()/qInfo(): swallow the stream unless GSB_REF_VERBOSE is set.
class GsbShimDebug
{
  public:
    GsbShimDebug() : m_on(std::getenv("GSB_REF_VERBOSE") != nullptr) {}
    GsbShimDebug(GsbShimDebug&& o) : m_on(o.m_on), m_ss(std::move(o.m_ss)) { o.m_on = false; }
    ~GsbShimDebug()
    {
        if (m_on)
            std::cerr << m_ss.str() << std::endl;
    }
    template <typename T> GsbShimDebug& operator<<(const T& v)
    {
        if (m_on)
            m_ss << v << ' ';
        return *this;
    }
    GsbShimDebug& operator<<(const QString& v)
    {
        if (m_on)
            m_ss << v.toStdString() << ' ';
        return *this;
    }

  private:
    bool m_on;
    std::ostringstream m_ss;
};
inline GsbShimDebug qDebug() { return GsbShimDebug(); }
inline GsbShimDebug qInfo() { return GsbShimDebug(); }

class QMutex
{
  public:
    void lock() { m_m.lock(); }
    void unlock() { m_m.unlock(); }

  private:
    std::mutex m_m;
};

template <typename T> class QFuture
{
  public:
    QFuture() = default;
    explicit QFuture(std::shared_future<T> f) : m_f(std::move(f)) {}
    void waitForFinished() { m_f.get(); } // rethrows, like QFuture does for QException

  private:
    std::shared_future<T> m_f;
};

class QThreadPool
{
};

namespace QtConcurrent
{
// QtConcurrent::run(obj, &Class::method, args...) -> runs on another thread (Qt: global pool).
template <typename C, typename M, typename... A>
QFuture<void> run(C* obj, M method, A... args)
{
    auto fut = std::async(std::launch::async,
                          [obj, method, args...]() { (obj->*method)(args...); });
    return QFuture<void>(fut.share());
}

// QtConcurrent::blockingMap(sequence, functor): apply functor to every element in place using
// the global pool (Qt: one thread per logical core).  GSB_REF_THREADS overrides the count.
template <typename Seq, typename F> void blockingMap(Seq& seq, F functor)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("GSB_REF_THREADS"))
        nt = static_cast<unsigned>(std::atoi(e));
    if (nt < 1)
        nt = 1;
    const size_t n = seq.size();
    if (nt == 1 || n < 4096) {
        for (auto& v : seq)
            functor(v);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
        if (lo == hi)
            break;
        pool.emplace_back([&seq, functor, lo, hi]() {
            for (size_t i = lo; i < hi; i++)
                functor(seq[i]);
        });
    }
    for (auto& th : pool)
        th.join();
}
} // namespace QtConcurrent
