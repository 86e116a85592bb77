// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal stand-ins for the Qt macros / typedefs that cbird's header-only search trees use, so that
// /root/reference/src/hamm.h, src/tree/vptree.h and src/tree/radix.h compile UNMODIFIED (included by
// path, never copied) without Qt.  Definitions mirror src/global.h:58-66 (typedefs) and Qt's public
// macro semantics; asserts stay enabled like the reference forces them (src/global.h:24-31).
#pragma once
#include <assert.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <cmath>
#include <vector>

typedef unsigned int uint;
typedef long long qint64;
typedef uint64_t dcthash_t;  // src/global.h:65
typedef uint32_t mediaid_t;  // src/global.h:66

#define Q_ASSERT(x) assert(x)
#define Q_LIKELY(x) __builtin_expect(!!(x), 1)
#define Q_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define Q_DISABLE_COPY_MOVE(C) \
  C(const C&) = delete;        \
  C& operator=(const C&) = delete; \
  C(C&&) = delete;             \
  C& operator=(C&&) = delete;

// qDebug()/qInfo()/qWarning(): both the printf form and the `<<` stream form appear in the headers
struct QtNullStream {
  template <typename T>
  QtNullStream& operator<<(const T&) { return *this; }
};
inline QtNullStream qt_quiet(const char* = nullptr, ...) { return QtNullStream(); }
inline QtNullStream qt_warn(const char* fmt = nullptr, ...) {
  if (fmt) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
  }
  return QtNullStream();
}
#define qWarning(...) qt_warn(__VA_ARGS__)
#define qInfo(...) qt_quiet(__VA_ARGS__)
#define qDebug(...) qt_quiet(__VA_ARGS__)

// src/global.h:58-62 (typed malloc/realloc helpers the tree headers call)
#define strict_malloc(ptr, count) reinterpret_cast<decltype(ptr)>(malloc(uint(count) * sizeof(*ptr)))
#define strict_realloc(ptr, count) reinterpret_cast<decltype(ptr)>(realloc(ptr, uint(count) * sizeof(*ptr)))
