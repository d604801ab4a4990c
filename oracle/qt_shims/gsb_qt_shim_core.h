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

#define QT_VERSION_CHECK(major, minor, patch) ((major << 16) | (minor << 8) | (patch))
#define QT_VERSION QT_VERSION_CHECK(5, 15, 0)
typedef unsigned long long quint64;

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
    bool isEmpty() const { return m_s.empty(); }
    // "%1".arg(x): replaces the lowest-numbered place marker
    QString arg(const QString& a) const
    {
        for (int n = 1; n < 10; n++) {
            const std::string marker = "%" + std::to_string(n);
            const size_t pos = m_s.find(marker);
            if (pos != std::string::npos) {
                std::string out = m_s;
                out.replace(pos, marker.size(), a.m_s);
                return QString(out);
            }
        }
        return *this;
    }
    int toInt(bool* ok = nullptr) const
    {
        char* end = nullptr;
        const long v = std::strtol(m_s.c_str(), &end, 10);
        const bool good = !m_s.empty() && end && *end == '\0';
        if (ok)
            *ok = good;
        return good ? static_cast<int>(v) : 0;
    }

  private:
    std::string m_s;
};
#define qPrintable(s) ((s).toStdString().c_str())
inline unsigned int qHash(const QString& s) { return static_cast<unsigned int>(std::hash<std::string>()(s.toStdString())); }
namespace std
{
template <> struct hash<QString> {
    std::size_t operator()(const QString& s) const { return std::hash<std::string>()(s.toStdString()); }
};
} // namespace std

class QStringList : public std::vector<QString>
{
  public:
    QStringList& operator<<(const QString& s)
    {
        push_back(s);
        return *this;
    }
};

// Parent / child ownership, sender() and pointer-to-member connect(): what the reference's
// GPUSimServer uses of QObject (gpusim.cpp:255-274, 294-304).  Signals are ordinary member
// functions of the stand-in classes; a class with signals implements gsbConnect().
#define slots
#define signals public
class QObject
{
  public:
    explicit QObject(QObject* parent = nullptr) : m_parent(parent)
    {
        if (parent)
            parent->m_children.push_back(this);
    }
    virtual ~QObject()
    {
        if (m_parent) {
            auto& c = m_parent->m_children;
            c.erase(std::remove(c.begin(), c.end(), this), c.end());
        }
        std::vector<QObject*> kids;
        kids.swap(m_children);
        for (QObject* k : kids) {
            k->m_parent = nullptr;
            delete k;
        }
    }
    QObject(const QObject&) = delete;
    QObject& operator=(const QObject&) = delete;
    QObject* sender() const { return m_sender; }
    void gsbSetSender(QObject* s) { m_sender = s; }
    void deleteLater() { m_delete_later = true; }
    bool gsbDeleteRequested() const { return m_delete_later; }
    template <class Snd, class Sig, class Rcv, class Slt>
    static bool connect(Snd* sender, Sig signal, Rcv* receiver, Slt slot)
    {
        return sender->gsbConnect(signal, [sender, receiver, slot]() {
            receiver->gsbSetSender(sender);
            (receiver->*slot)();
            receiver->gsbSetSender(nullptr);
        });
    }

  private:
    QObject* m_parent = nullptr;
    QObject* m_sender = nullptr;
    bool m_delete_later = false;
    std::vector<QObject*> m_children;
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
inline GsbShimDebug qWarning() { return GsbShimDebug(); }

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

// QRunnable / QThreadPool as GPUSimServer::extractData uses them (gpusim.cpp:48-85, 202-236):
// start() runs the job on its own thread and deletes it (autoDelete), waitForDone() joins.
class QRunnable
{
  public:
    virtual ~QRunnable() = default;
    virtual void run() = 0;
};
class QThreadPool
{
  public:
    ~QThreadPool() { waitForDone(); }
    void start(QRunnable* job)
    {
        // never more jobs in flight than cores (Qt: maxThreadCount = ideal thread count)
        unsigned n = std::thread::hardware_concurrency();
        if (m_threads.size() >= (n ? n : 1u))
            waitForDone();
        m_threads.emplace_back([job]() {
            job->run();
            delete job;
        });
    }
    void waitForDone()
    {
        for (auto& t : m_threads)
            t.join();
        m_threads.clear();
    }

  private:
    std::vector<std::thread> m_threads;
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
