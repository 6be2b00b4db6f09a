// conversion.hpp -- iqs::toString (interface of reference include/conversion.hpp:34-40).
#ifndef IQS_CONVERSION_HPP
#define IQS_CONVERSION_HPP
#include <cassert>
#include <sstream>
#include <string>
namespace iqs {
template <class T>
std::string toString(T const &val) {
  std::ostringstream os;
  os << val;
  return os.str();
}
}  // namespace iqs
#endif
