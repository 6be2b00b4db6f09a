// qureg_version.hpp -- version string of the library (interface of reference include/qureg_version.hpp:17-25).
#pragma once
#include <string>

#define QHIPSTER_VERSION_STRING "SDK-RC-2.1.0"
#define IQS_B200_ENGINE_VERSION "b200-r1"

namespace iqs {
std::string GetQhipsterVersion(void);
}  // namespace iqs
