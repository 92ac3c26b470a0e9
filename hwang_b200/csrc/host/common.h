// Types that cross the plugin boundary.  Mirrors hwang/common.h:20-68 of the reference
// (DeviceType, DeviceHandle, CPU_DEVICE, Result, HWANG_RETURN_ON_ERROR) without the glog dependency.
#pragma once
#include <stdint.h>
#include <string>

namespace hwang {

enum class DeviceType {
  CPU = 0,
  GPU = 1,
};

struct DeviceHandle {
  bool operator==(const DeviceHandle &o) const { return type == o.type && id == o.id; }
  bool operator!=(const DeviceHandle &o) const { return !(*this == o); }
  // strict weak order (type, then id), so that DeviceHandle can key a std::map; the reference's version
  // (hwang/common.h:33-35: `type < o.type && id < o.id`) is not one: {CPU,1} and {GPU,0} compare equivalent both ways
  bool operator<(const DeviceHandle &o) const { return type != o.type ? type < o.type : id < o.id; }
  bool can_copy_to(const DeviceHandle &o) const {
    return !(type == DeviceType::GPU && o.type == DeviceType::GPU && id != o.id);
  }
  bool is_same_address_space(const DeviceHandle &o) const {
    return type == o.type && (type == DeviceType::CPU || (type == DeviceType::GPU && id == o.id));
  }
  DeviceType type;
  int32_t id;
};

static const DeviceHandle CPU_DEVICE = {DeviceType::CPU, 0};

struct Result {
  Result() : ok(true) {}
  Result(bool _ok, const std::string &_message) : ok(_ok), message(_message) {}
  bool ok;
  std::string message;
};

#define HWANG_RETURN_ON_ERROR(expr__) \
  do {                                \
    ::hwang::Result res__ = (expr__); \
    if (!res__.ok) return res__;      \
  } while (0)

}  // namespace hwang
