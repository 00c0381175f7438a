// ompl_min: console/log macros (OMPL_ERROR / OMPL_WARN / OMPL_INFORM / OMPL_DEBUG).
#ifndef OMPL_MIN_UTIL_CONSOLE_H
#define OMPL_MIN_UTIL_CONSOLE_H
#include <cstdarg>
#include <cstdio>
#include <string>
namespace ompl {
namespace msg {
enum LogLevel { LOG_DEBUG = 0, LOG_INFO, LOG_WARN, LOG_ERROR, LOG_NONE };

class OutputHandler {
public:
  virtual ~OutputHandler() {}
  virtual void log(const std::string &text, LogLevel level, const char *filename, int line) = 0;
};

class OutputHandlerSTD : public OutputHandler {
public:
  void log(const std::string &text, LogLevel level, const char *, int) override {
    static const char *tag[] = {"Debug:   ", "Info:    ", "Warning: ", "Error:   ", ""};
    std::FILE *f = level >= LOG_WARN ? stderr : stdout;
    std::fprintf(f, "%s%s\n", tag[level], text.c_str());
  }
};

struct ConsoleState {
  LogLevel level = LOG_WARN;
  OutputHandler *handler = nullptr;
  OutputHandlerSTD std_handler;
  ConsoleState() { handler = &std_handler; }
};
inline ConsoleState &consoleState() {
  static ConsoleState s;
  return s;
}
inline void setLogLevel(LogLevel level) { consoleState().level = level; }
inline LogLevel getLogLevel() { return consoleState().level; }
inline void useOutputHandler(OutputHandler *oh) { consoleState().handler = oh; }
inline void noOutputHandler() { consoleState().handler = nullptr; }
inline OutputHandler *getOutputHandler() { return consoleState().handler; }

inline void log(const char *file, int line, LogLevel level, const char *fmt, ...)
    __attribute__((format(printf, 4, 5)));
inline void log(const char *file, int line, LogLevel level, const char *fmt, ...) {
  ConsoleState &s = consoleState();
  if (!s.handler || level < s.level)
    return;
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  s.handler->log(buf, level, file, line);
}
} // namespace msg
} // namespace ompl

#define OMPL_ERROR(fmt, ...) ompl::msg::log(__FILE__, __LINE__, ompl::msg::LOG_ERROR, fmt, ##__VA_ARGS__)
#define OMPL_WARN(fmt, ...) ompl::msg::log(__FILE__, __LINE__, ompl::msg::LOG_WARN, fmt, ##__VA_ARGS__)
#define OMPL_INFORM(fmt, ...) ompl::msg::log(__FILE__, __LINE__, ompl::msg::LOG_INFO, fmt, ##__VA_ARGS__)
#define OMPL_DEBUG(fmt, ...) ompl::msg::log(__FILE__, __LINE__, ompl::msg::LOG_DEBUG, fmt, ##__VA_ARGS__)
#endif
