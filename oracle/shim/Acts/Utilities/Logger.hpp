#pragma once
// Stand-in for Acts::Logger: the reference's algorithms only clone loggers and stream into
// level macros; nothing is printed here.
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
namespace Acts {
namespace Logging {
enum Level { VERBOSE = 0, DEBUG, INFO, WARNING, ERROR, FATAL, MAX };
}
class Logger {
    public:
    std::unique_ptr<Logger> clone() const { return std::make_unique<Logger>(); }
    std::unique_ptr<Logger> clone(const std::string&) const { return std::make_unique<Logger>(); }
    std::unique_ptr<Logger> cloneWithSuffix(const std::string&) const { return std::make_unique<Logger>(); }
    bool doPrint(Logging::Level) const { return false; }
    const std::string& name() const { static const std::string n = "shim"; return n; }
    Logging::Level level() const { return Logging::FATAL; }
};
inline const Logger& getDummyLogger() {
    static const Logger l;
    return l;
}
inline std::unique_ptr<const Logger> getDefaultLogger(const std::string&, Logging::Level) {
    return std::make_unique<const Logger>();
}
}  // namespace Acts
#define ACTS_LOCAL_LOGGER(x)
#define ACTS_LOG(level, x) do { } while (0)
#define ACTS_VERBOSE(x) do { } while (0)
#define ACTS_DEBUG(x) do { } while (0)
#define ACTS_INFO(x) do { } while (0)
#define ACTS_WARNING(x) do { } while (0)
#define ACTS_ERROR(x) do { } while (0)
#define ACTS_FATAL(x) do { } while (0)
