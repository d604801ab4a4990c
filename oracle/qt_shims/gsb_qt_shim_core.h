// Header-only stand-ins for the handful of Qt types the reference's hot-path sources touch.
// TEST INFRASTRUCTURE: lets oracle/Makefile compile /root/reference/{calculation_functors.cpp,
// fingerprintdb_cuda.cpp,fingerprintdb_cuda.cu} verbatim (Qt5 is not installed in this image)
// so the reference itself can serve as parity checker and CPU/CUDA baseline.  Not Qt, not
// part of the product; only the members those three files use exist.
#pragma once
#include <algorithm>
#include <climits>
#include <condition_variable>
#include <memory>
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
// The global thread pool Qt hands QtConcurrent::run jobs to: persistent workers (one per logical
// core, like QThreadPool::globalInstance()), so that per-thread CUDA state is set up once and not
// on every query as it would be with a fresh std::thread per job.
class GsbShimPool
{
  public:
    static GsbShimPool& instance()
    {
        // leaked on purpose: the workers block on the condition variable for the life of the
        // process, and destroying a condition variable with waiters at exit would hang
        static GsbShimPool* pool = new GsbShimPool;
        return *pool;
    }
    std::shared_future<void> submit(std::function<void()> fn)
    {
        auto task = std::make_shared<std::packaged_task<void()>>(std::move(fn));
        std::shared_future<void> fut = task->get_future().share();
        {
            std::lock_guard<std::mutex> lock(m_mu);
            m_jobs.push_back([task]() { (*task)(); });
        }
        m_cv.notify_one();
        return fut;
    }

  private:
    GsbShimPool()
    {
        unsigned n = std::thread::hardware_concurrency();
        if (n < 1)
            n = 1;
        for (unsigned i = 0; i < n; i++)
            std::thread([this]() { work(); }).detach();
    }
    void work()
    {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lock(m_mu);
                m_cv.wait(lock, [this]() { return !m_jobs.empty(); });
                job = std::move(m_jobs.front());
                m_jobs.erase(m_jobs.begin());
            }
            job();
        }
    }
    std::mutex m_mu;
    std::condition_variable m_cv;
    std::vector<std::function<void()>> m_jobs;
};

// QtConcurrent::run(obj, &Class::method, args...) -> runs on the global pool.
template <typename C, typename M, typename... A>
QFuture<void> run(C* obj, M method, A... args)
{
    return QFuture<void>(
        GsbShimPool::instance().submit([obj, method, args...]() { (obj->*method)(args...); }));
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
