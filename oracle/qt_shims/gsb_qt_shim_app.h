// Stand-ins for the application-level Qt classes the reference's gpusim.cpp, main.cpp and
// test/test_gpusim.cpp use: QByteArray, qUncompress, QIODevice, QFile, QFileInfo, QDataStream,
// QHash, QLocalServer, QLocalSocket, QCoreApplication, QCommandLineParser, QElapsedTimer.
// TEST INFRASTRUCTURE (Qt5 is not installed in this image): `make refcheck` compiles those three
// reference files UNMODIFIED against include/gpusim/ + these headers and links them with
// libgpusim_adapter.so, which proves the drop-in boundary and lets the reference's own test-suite
// and daemon run on top of the B200 engine.  Functional, not Qt: only what those files touch.
//   QDataStream   big-endian; char* = u32 length (with NUL) + bytes, new[]-allocated on read;
//                 QByteArray = u32 length + bytes (0xFFFFFFFF = null); float travels as an 8-byte
//                 double (Qt >= 4.6 default DoublePrecision) — SURVEY App. A / B
//   qUncompress   4-byte big-endian length + zlib stream
//   QLocalServer  unix stream socket at /tmp/<name>; QCoreApplication::exec() is a poll() loop that
//                 raises newConnection / readyRead / disconnected
#pragma once
#include "gsb_qt_shim_core.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>
#include <zlib.h>

class QByteArray
{
  public:
    QByteArray() = default;
    QByteArray(const char* d, int n) : m_d(d, d + n) {}
    char* data() { return m_d.data(); }
    const char* data() const { return m_d.data(); }
    const char* constData() const { return m_d.data(); }
    int size() const { return static_cast<int>(m_d.size()); }
    bool isEmpty() const { return m_d.empty(); }
    void clear()
    {
        m_d.clear();
        m_d.shrink_to_fit();
    }
    void append(const char* d, size_t n) { m_d.insert(m_d.end(), d, d + n); }
    void resize(size_t n) { m_d.resize(n); }

  private:
    std::vector<char> m_d;
};

inline QByteArray qUncompress(const QByteArray& in)
{
    QByteArray out;
    if (in.size() < 4)
        return out;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(in.constData());
    uLongf len = (uLongf(p[0]) << 24) | (uLongf(p[1]) << 16) | (uLongf(p[2]) << 8) | uLongf(p[3]);
    out.resize(len ? len : 1);
    if (uncompress(reinterpret_cast<Bytef*>(out.data()), &len, p + 4, in.size() - 4) != Z_OK)
        return QByteArray();
    out.resize(len);
    return out;
}

class QIODevice
{
  public:
    enum OpenModeFlag { NotOpen = 0, ReadOnly = 1, WriteOnly = 2, ReadWrite = 3 };
    virtual ~QIODevice() = default;
    virtual size_t gsbRead(char* dst, size_t n) = 0;
    virtual bool atEnd() = 0;
};

class QFile : public QIODevice
{
  public:
    explicit QFile(const QString& name) : m_name(name.toStdString()) {}
    ~QFile() override
    {
        if (m_f)
            std::fclose(m_f);
    }
    bool open(int)
    {
        m_f = std::fopen(m_name.c_str(), "rb");
        return m_f != nullptr;
    }
    static bool remove(const QString& name) { return ::unlink(name.toStdString().c_str()) == 0; }
    size_t gsbRead(char* dst, size_t n) override { return m_f ? std::fread(dst, 1, n, m_f) : 0; }
    bool atEnd() override
    {
        if (!m_f)
            return true;
        const int c = std::fgetc(m_f);
        if (c == EOF)
            return true;
        std::ungetc(c, m_f);
        return false;
    }

  private:
    std::string m_name;
    FILE* m_f = nullptr;
};

class QFileInfo
{
  public:
    explicit QFileInfo(const QString& name) : m_name(name.toStdString()) {}
    bool exists() const
    {
        struct stat st;
        return ::stat(m_name.c_str(), &st) == 0;
    }
    QString baseName() const // file name without path, up to the first '.'
    {
        const size_t slash = m_name.find_last_of('/');
        std::string base = slash == std::string::npos ? m_name : m_name.substr(slash + 1);
        const size_t dot = base.find('.');
        return QString(dot == std::string::npos ? base : base.substr(0, dot));
    }

  private:
    std::string m_name;
};

class QDataStream
{
  public:
    enum Version { Qt_5_2 = 15 };
    explicit QDataStream(QIODevice* dev) : m_dev(dev) {}
    QDataStream(QByteArray* ba, int mode) : m_ba(ba), m_write(mode == QIODevice::WriteOnly) {}
    explicit QDataStream(const QByteArray& ba) : m_ba(const_cast<QByteArray*>(&ba)) {}
    void setVersion(int) {}
    bool atEnd()
    {
        if (m_dev)
            return m_dev->atEnd();
        return m_pos >= static_cast<size_t>(m_ba->size());
    }
    QDataStream& operator>>(int& v)
    {
        unsigned char b[4] = {0, 0, 0, 0};
        get(b, 4);
        v = static_cast<int>((uint32_t(b[0]) << 24) | (uint32_t(b[1]) << 16) | (uint32_t(b[2]) << 8) | uint32_t(b[3]));
        return *this;
    }
    QDataStream& operator>>(float& v) // double on the wire
    {
        unsigned char b[8] = {0};
        get(b, 8);
        uint64_t u = 0;
        for (int i = 0; i < 8; i++)
            u = (u << 8) | b[i];
        double d;
        std::memcpy(&d, &u, 8);
        v = static_cast<float>(d);
        return *this;
    }
    QDataStream& operator>>(char*& s) // new[]-allocated, like Qt
    {
        int len = 0;
        *this >> len;
        if (len <= 0) {
            s = new char[1];
            s[0] = '\0';
            return *this;
        }
        s = new char[static_cast<size_t>(len) + 1];
        get(reinterpret_cast<unsigned char*>(s), static_cast<size_t>(len));
        s[len] = '\0';
        return *this;
    }
    QDataStream& operator>>(QByteArray& ba)
    {
        int len = 0;
        *this >> len;
        ba.clear();
        if (len > 0) {
            ba.resize(static_cast<size_t>(len));
            get(reinterpret_cast<unsigned char*>(ba.data()), static_cast<size_t>(len));
        }
        return *this;
    }
    QDataStream& operator<<(int v)
    {
        const uint32_t u = static_cast<uint32_t>(v);
        const unsigned char b[4] = {static_cast<unsigned char>(u >> 24), static_cast<unsigned char>(u >> 16),
                                    static_cast<unsigned char>(u >> 8), static_cast<unsigned char>(u)};
        put(b, 4);
        return *this;
    }
    QDataStream& operator<<(quint64 u)
    {
        unsigned char b[8];
        for (int i = 0; i < 8; i++)
            b[i] = static_cast<unsigned char>(u >> (56 - 8 * i));
        put(b, 8);
        return *this;
    }
    QDataStream& operator<<(float v) // double on the wire
    {
        const double d = v;
        uint64_t u;
        std::memcpy(&u, &d, 8);
        return *this << static_cast<quint64>(u);
    }
    QDataStream& operator<<(const char* s)
    {
        if (!s)
            return *this << static_cast<int>(0xffffffffu);
        const size_t len = std::strlen(s) + 1;
        *this << static_cast<int>(len);
        put(reinterpret_cast<const unsigned char*>(s), len);
        return *this;
    }

  private:
    void get(unsigned char* dst, size_t n)
    {
        if (m_dev) {
            m_dev->gsbRead(reinterpret_cast<char*>(dst), n);
            return;
        }
        const size_t have = static_cast<size_t>(m_ba->size());
        const size_t take = m_pos + n <= have ? n : (m_pos < have ? have - m_pos : 0);
        std::memcpy(dst, m_ba->constData() + m_pos, take);
        m_pos += take;
    }
    void put(const unsigned char* src, size_t n)
    {
        if (m_ba && m_write)
            m_ba->append(reinterpret_cast<const char*>(src), n);
    }
    QIODevice* m_dev = nullptr;
    QByteArray* m_ba = nullptr;
    bool m_write = false;
    size_t m_pos = 0;
};

// QHash as the reference uses it: operator[], contains(), and range-for over the VALUES.
template <class K, class V> class QHash
{
    typedef std::map<K, V> Map;

  public:
    class iterator
    {
      public:
        explicit iterator(typename Map::iterator it) : m_it(it) {}
        V& operator*() const { return m_it->second; }
        iterator& operator++()
        {
            ++m_it;
            return *this;
        }
        bool operator!=(const iterator& o) const { return m_it != o.m_it; }

      private:
        typename Map::iterator m_it;
    };
    V& operator[](const K& k) { return m_map[k]; }
    bool contains(const K& k) const { return m_map.count(k) > 0; }
    iterator begin() { return iterator(m_map.begin()); }
    iterator end() { return iterator(m_map.end()); }
    int size() const { return static_cast<int>(m_map.size()); }

  private:
    Map m_map;
};

class QElapsedTimer
{
  public:
    void start() { m_t0 = std::chrono::steady_clock::now(); }
    long long elapsed() const
    {
        return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - m_t0).count();
    }

  private:
    std::chrono::steady_clock::time_point m_t0 = std::chrono::steady_clock::now();
};
class QTime
{
};
class QThread
{
};
class QSize
{
};

// ---- local sockets + event loop ------------------------------------------------------------
class QLocalServer;
class QLocalSocket;
struct GsbShimLoop {
    std::vector<QLocalServer*> servers;
    std::vector<QLocalSocket*> sockets;
    bool quit = false;
    int code = 0;
    static GsbShimLoop& instance()
    {
        static GsbShimLoop loop;
        return loop;
    }
};

class QLocalSocket : public QObject
{
  public:
    explicit QLocalSocket(int fd) : m_fd(fd) { GsbShimLoop::instance().sockets.push_back(this); }
    ~QLocalSocket() override
    {
        auto& v = GsbShimLoop::instance().sockets;
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
        if (m_fd >= 0)
            ::close(m_fd);
    }
    QByteArray readAll()
    {
        QByteArray out;
        out.append(m_in.data(), m_in.size());
        m_in.clear();
        return out;
    }
    long long write(const QByteArray& d)
    {
        size_t off = 0;
        while (off < static_cast<size_t>(d.size())) {
            const ssize_t w = ::send(m_fd, d.constData() + off, d.size() - off, MSG_NOSIGNAL);
            if (w <= 0)
                return -1;
            off += static_cast<size_t>(w);
        }
        return static_cast<long long>(off);
    }
    bool flush() { return true; }
    // signals
    void disconnected() { gsbEmit(m_disconnected); }
    void readyRead() { gsbEmit(m_ready_read); }
    bool gsbConnect(void (QLocalSocket::*sig)(), std::function<void()> fn)
    {
        if (sig == &QLocalSocket::disconnected)
            m_disconnected.push_back(std::move(fn));
        else if (sig == &QLocalSocket::readyRead)
            m_ready_read.push_back(std::move(fn));
        else
            return false;
        return true;
    }
    int gsbFd() const { return m_fd; }
    std::vector<char>& gsbInput() { return m_in; }

  private:
    static void gsbEmit(const std::vector<std::function<void()>>& slots_)
    {
        for (const auto& fn : slots_)
            fn();
    }
    int m_fd;
    std::vector<char> m_in;
    std::vector<std::function<void()>> m_disconnected, m_ready_read;
};

class QLocalServer : public QObject
{
  public:
    explicit QLocalServer(QObject* parent = nullptr) : QObject(parent) {}
    ~QLocalServer() override { close(); }
    bool listen(const QString& name)
    {
        const std::string n = name.toStdString();
        m_path = !n.empty() && n[0] == '/' ? n : "/tmp/" + n; // QDir::tempPath() + name
        m_fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
        if (m_fd < 0)
            return false;
        sockaddr_un addr;
        std::memset(&addr, 0, sizeof(addr));
        addr.sun_family = AF_UNIX;
        std::strncpy(addr.sun_path, m_path.c_str(), sizeof(addr.sun_path) - 1);
        if (::bind(m_fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || ::listen(m_fd, 16) != 0) {
            ::close(m_fd); // an existing file makes bind fail, as with Qt
            m_fd = -1;
            return false;
        }
        m_bound = true;
        GsbShimLoop::instance().servers.push_back(this);
        return true;
    }
    void close()
    {
        auto& v = GsbShimLoop::instance().servers;
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
        if (m_fd >= 0)
            ::close(m_fd);
        m_fd = -1;
        if (m_bound)
            ::unlink(m_path.c_str());
        m_bound = false;
    }
    QLocalSocket* nextPendingConnection()
    {
        if (m_pending.empty())
            return nullptr;
        QLocalSocket* s = m_pending.front();
        m_pending.erase(m_pending.begin());
        return s;
    }
    // signal
    void newConnection()
    {
        for (const auto& fn : m_new_connection)
            fn();
    }
    bool gsbConnect(void (QLocalServer::*sig)(), std::function<void()> fn)
    {
        if (sig != &QLocalServer::newConnection)
            return false;
        m_new_connection.push_back(std::move(fn));
        return true;
    }
    int gsbFd() const { return m_fd; }
    void gsbAccepted(QLocalSocket* s) { m_pending.push_back(s); }

  private:
    int m_fd = -1;
    bool m_bound = false;
    std::string m_path;
    std::vector<QLocalSocket*> m_pending;
    std::vector<std::function<void()>> m_new_connection;
};

class QCoreApplication : public QObject
{
  public:
    QCoreApplication(int& argc, char** argv)
    {
        gsbArgs().clear();
        for (int i = 0; i < argc; i++)
            gsbArgs() << QString(argv[i]);
        gsbInstance() = this;
    }
    ~QCoreApplication() override { gsbInstance() = nullptr; }
    static void setApplicationName(const QString&) {}
    static QStringList arguments() { return gsbArgs(); }
    static QCoreApplication* instance() { return gsbInstance(); }
    static void exit(int code = 0)
    {
        GsbShimLoop::instance().quit = true;
        GsbShimLoop::instance().code = code;
    }
    static void quit() { exit(0); }
    // poll() loop: accept -> newConnection, data -> readyRead, EOF -> disconnected (+ deleteLater)
    static int exec()
    {
        GsbShimLoop& loop = GsbShimLoop::instance();
        while (!loop.quit) {
            std::vector<pollfd> fds;
            const std::vector<QLocalServer*> servers = loop.servers;
            const std::vector<QLocalSocket*> sockets = loop.sockets;
            for (QLocalServer* s : servers)
                fds.push_back({s->gsbFd(), POLLIN, 0});
            for (QLocalSocket* s : sockets)
                fds.push_back({s->gsbFd(), POLLIN, 0});
            if (::poll(fds.data(), fds.size(), 100) <= 0)
                continue;
            size_t i = 0;
            for (QLocalServer* s : servers) {
                if (fds[i++].revents & POLLIN) {
                    const int fd = ::accept(s->gsbFd(), nullptr, nullptr);
                    if (fd >= 0) {
                        s->gsbAccepted(new QLocalSocket(fd));
                        s->newConnection();
                    }
                }
            }
            for (QLocalSocket* s : sockets) {
                const short re = fds[i++].revents;
                if (!(re & (POLLIN | POLLHUP | POLLERR)))
                    continue;
                char buf[65536];
                const ssize_t n = ::recv(s->gsbFd(), buf, sizeof(buf), 0);
                if (n > 0) {
                    s->gsbInput().insert(s->gsbInput().end(), buf, buf + n);
                    // one readAll() per request is assumed by the reference: drain what
                    // has already arrived before raising the signal
                    for (;;) {
                        pollfd more = {s->gsbFd(), POLLIN, 0};
                        if (::poll(&more, 1, 2) <= 0 || !(more.revents & POLLIN))
                            break;
                        const ssize_t m = ::recv(s->gsbFd(), buf, sizeof(buf), 0);
                        if (m <= 0)
                            break;
                        s->gsbInput().insert(s->gsbInput().end(), buf, buf + m);
                    }
                    s->readyRead();
                } else {
                    s->disconnected();
                    if (s->gsbDeleteRequested())
                        delete s;
                    else
                        ::shutdown(s->gsbFd(), SHUT_RDWR);
                }
            }
        }
        return loop.code;
    }

  private:
    static QStringList& gsbArgs()
    {
        static QStringList args;
        return args;
    }
    static QCoreApplication*& gsbInstance()
    {
        static QCoreApplication* app = nullptr;
        return app;
    }
};

// ---- command line (main.cpp:16-61) -----------------------------------------------------------
class QCommandLineOption
{
  public:
    QCommandLineOption(const QString& name, const QString& description = QString(), const QString& value_name = QString(),
                       const QString& default_value = QString())
        : m_name(name.toStdString()), m_takes_value(!value_name.isEmpty()), m_default(default_value.toStdString())
    {
        (void) description;
    }
    std::string m_name;
    bool m_takes_value;
    std::string m_default;
};

class QCommandLineParser
{
  public:
    void setApplicationDescription(const QString&) {}
    QCommandLineOption addHelpOption()
    {
        QCommandLineOption help("help");
        m_options.push_back(help);
        return help;
    }
    bool addOption(const QCommandLineOption& o)
    {
        m_options.push_back(o);
        return true;
    }
    bool parse(const QStringList& args)
    {
        for (size_t i = 1; i < args.size(); i++) {
            const std::string a = args[i].toStdString();
            if (a.size() < 2 || a[0] != '-') {
                m_positional << args[i];
                continue;
            }
            std::string name = a.substr(a[1] == '-' ? 2 : 1), value;
            bool has_value = false;
            const size_t eq = name.find('=');
            if (eq != std::string::npos) {
                value = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_value = true;
            }
            if (name == "h")
                name = "help";
            const QCommandLineOption* opt = nullptr;
            for (const auto& o : m_options)
                if (o.m_name == name)
                    opt = &o;
            if (!opt) {
                m_error = "Unknown option '" + name + "'.";
                return false;
            }
            if (opt->m_takes_value && !has_value) {
                if (i + 1 >= args.size()) {
                    m_error = "Missing value after '" + a + "'.";
                    return false;
                }
                value = args[++i].toStdString();
            }
            m_values[name] = value;
        }
        return true;
    }
    QString errorText() const { return QString(m_error); }
    bool isSet(const QCommandLineOption& o) const { return m_values.count(o.m_name) > 0; }
    QString value(const QCommandLineOption& o) const
    {
        const auto it = m_values.find(o.m_name);
        return QString(it == m_values.end() ? o.m_default : it->second);
    }
    QStringList positionalArguments() const { return m_positional; }

  private:
    std::vector<QCommandLineOption> m_options;
    std::map<std::string, std::string> m_values;
    QStringList m_positional;
    std::string m_error;
};
