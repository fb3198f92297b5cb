// logger.h -- Boost-free logger with the reference's spelling (reference src/logger.h:32-85):
//   log_msg<LOG_INFO>(L"text %d") % value;     c_log_msg(LOG_DEBUG, "printf style %d", v);
// Messages at level <= LOG_TO_FILE go to solver_log.txt, at level <= LOG_COUT to stdout, like the
// reference; the file is opened once (the reference re-opens it per message) and nothing in the
// step loop logs.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <sstream>
#include <string>

enum log_level_t { LOG_NOTHING = 0, LOG_CRITICAL, LOG_ERROR, LOG_WARNING, LOG_INFO, LOG_CONFIG, LOG_DEBUG, LOG_VERBOSE, LOG_TRACE };

#ifndef LOG_TO_FILE
#define LOG_TO_FILE 4
#endif
#ifndef LOG_COUT
#define LOG_COUT 2
#endif

namespace pfdtd_host {
inline FILE*& log_file() { static FILE* f = nullptr; return f; }
inline std::mutex& log_mutex() { static std::mutex m; return m; }
inline void log_emit(int level, const std::string& msg) {
  std::lock_guard<std::mutex> g(log_mutex());
  if (level <= LOG_TO_FILE) {
    if (!log_file()) log_file() = std::fopen("solver_log.txt", "a");
    if (log_file()) { std::fprintf(log_file(), "%d %s\n", level, msg.c_str()); std::fflush(log_file()); }
  }
  if (level <= LOG_COUT) std::printf("%d %s\n", level, msg.c_str());
}
inline std::string narrow(const wchar_t* w) {
  std::string s;
  for (; w && *w; ++w) s.push_back(*w < 128 ? (char)*w : '?');
  return s;
}
// minimal positional formatter: every `% arg` replaces the next printf-like directive
class LogLine {
 public:
  LogLine(int level, const wchar_t* fmt) : level_(level), fmt_(narrow(fmt)), pos_(0) {}
  LogLine(const LogLine&) = delete;
  LogLine(LogLine&& o) : level_(o.level_), fmt_(std::move(o.fmt_)), out_(std::move(o.out_)), pos_(o.pos_) { o.level_ = -1; }
  ~LogLine() { if (level_ >= 0) { out_ += fmt_.substr(pos_); log_emit(level_, out_); } }
  template <typename T> LogLine& operator%(const T& v) {
    size_t p = fmt_.find('%', pos_);
    if (p == std::string::npos) { std::ostringstream o; o << " [" << v << "]"; tail_(o.str()); return *this; }
    out_ += fmt_.substr(pos_, p - pos_);
    size_t e = p + 1;
    while (e < fmt_.size() && std::string("diouxXeEfFgGscp").find(fmt_[e]) == std::string::npos) e++;
    std::ostringstream o; o << v; out_ += o.str();
    pos_ = e < fmt_.size() ? e + 1 : fmt_.size();
    return *this;
  }
 private:
  void tail_(const std::string& s) { fmt_ += s; }
  int level_; std::string fmt_, out_; size_t pos_;
};
}  // namespace pfdtd_host

inline void loggerInit() { std::lock_guard<std::mutex> g(pfdtd_host::log_mutex()); FILE* f = std::fopen("solver_log.txt", "w"); if (f) std::fclose(f); }
template <log_level_t LEVEL> inline pfdtd_host::LogLine log_msg(const wchar_t* fmt) { return pfdtd_host::LogLine((int)LEVEL, fmt); }
inline void c_log_msg(log_level_t level, const char* fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  pfdtd_host::log_emit((int)level, buf);
}
