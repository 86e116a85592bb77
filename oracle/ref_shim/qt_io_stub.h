// TEST INFRASTRUCTURE ONLY. Working stand-ins (over std::string / FILE*) for the few Qt value and IO
// classes that /root/reference/src/tree/hammingtree.h uses for its cache-file reader/writer
// (:156-200, :456-521), so that header compiles unmodified AND its read()/write() actually run.
#pragma once
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <type_traits>
#include <unordered_set>
#include <vector>

#include "qt_shim.h"

template <typename T>
class QList : public std::vector<T> {
 public:
  int count() const { return int(this->size()); }
};

class QString;

class QByteArray {
 public:
  std::string s;
  QByteArray() {}
  QByteArray(const char* p) : s(p) {}
  QByteArray(const std::string& x) : s(x) {}
  int length() const { return int(s.size()); }
  void resize(int n) { s.resize(size_t(n)); }
  const char* data() const { return s.data(); }
  QByteArray& append(const char* p, size_t n) {
    s.append(p, n);
    return *this;
  }
  QList<QByteArray> split(char sep) const {
    QList<QByteArray> out;
    size_t start = 0;
    for (;;) {
      size_t p = s.find(sep, start);
      if (p == std::string::npos) {
        out.push_back(QByteArray(s.substr(start)));
        break;
      }
      out.push_back(QByteArray(s.substr(start, p - start)));
      start = p + 1;
    }
    return out;
  }
  bool operator!=(const char* o) const { return s != o; }
  bool operator!=(const QString& o) const;
};

class QString {
 public:
  std::string s;
  QString() {}
  QString(const char* p) : s(p) {}
  QString(const std::string& x) : s(x) {}
  QString(const QByteArray& b) : s(b.s) {}
  template <typename T>
  QString arg(T v) const {  // replaces the lowest-numbered %N marker
    int lowest = 100;
    for (size_t i = 0; i + 1 < s.size(); ++i)
      if (s[i] == '%' && s[i + 1] >= '1' && s[i + 1] <= '9') lowest = std::min(lowest, s[i + 1] - '0');
    std::string out = s, marker = "%" + std::to_string(lowest), val = std::to_string((long long)v);
    for (size_t p = out.find(marker); p != std::string::npos; p = out.find(marker, p + val.size()))
      out.replace(p, marker.size(), val);
    return QString(out);
  }
  static QString number(long long v) { return QString(std::to_string(v)); }
  QByteArray toLatin1() const { return QByteArray(s); }
};
inline bool QByteArray::operator!=(const QString& o) const { return s != o.s; }
#define QStringLiteral(x) QString(x)

class QFile {
 public:
  FILE* fp = nullptr;
  explicit QFile(FILE* f) : fp(f) {}
  QByteArray readLine(int maxlen) {
    std::string line;
    int c;
    while (int(line.size()) < maxlen - 1 && (c = fgetc(fp)) != EOF) {
      line.push_back(char(c));
      if (c == '\n') break;
    }
    return QByteArray(line);
  }
  qint64 read(char* dst, qint64 len) { return qint64(fread(dst, 1, size_t(len), fp)); }
  qint64 write(const char* p) { return qint64(fwrite(p, 1, strlen(p), fp)); }
  qint64 write(const QByteArray& b) { return qint64(fwrite(b.s.data(), 1, b.s.size(), fp)); }
  bool atEnd() {
    int c = fgetc(fp);
    if (c == EOF) return true;
    ungetc(c, fp);
    return false;
  }
  QString errorString() const { return QString("io error"); }
};
