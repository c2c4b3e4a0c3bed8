// sb200/core.h -- foundations of the C++ host layer: exceptions, type identity, parameters,
// contexts and the glue that turns C-ABI error codes (include/sb200.h) into the exceptions a
// SparseBase user expects.
//
// The host layer re-creates, from scratch and only for the preprocessing hot path, the
// public surface of sparcityeu/SparseBase v0.3.1 (namespace sparsebase::{utils, context,
// format, converter, reorder, permute, feature, bases}) so that user code written against
// the reference compiles against these headers and runs on a B200 through libsb200.so.
// Nothing here computes on the CPU: every operator ends in a C-ABI call.
//
// Interfaces mirrored in this file (reference paths relative to src/sparsebase/):
//   utils::Exception + subclasses      utils/exception.h:23-201
//   utils::Identifiable[Implementation] utils/utils.h:176-199
//   utils::Parameters / Parameterizable utils/parameterizable.h:11-21
//   context::Context                   context/context.h:18-21
//   context::CPUContext                context/cpu_context.h:12-14
//   context::CUDAContext               context/cuda_context_cuda.cuh:15-19, .cu:9-22
#pragma once
#include <cxxabi.h>

#include <cstdint>
#include <cstdlib>
#include <exception>
#include <memory>
#include <string>
#include <type_traits>
#include <typeindex>
#include <typeinfo>
#include <vector>

#include "../../../../include/sb200.h"

namespace sparsebase {

// ======================================================================= utils
namespace utils {

typedef unsigned int CostType;

template <typename T>
inline constexpr bool always_false = false;

class Exception : public std::exception {};

namespace detail {
class MessageException : public Exception {
 public:
  explicit MessageException(std::string msg) : msg_(std::move(msg)) {}
  const char *what() const noexcept override { return msg_.c_str(); }

 protected:
  std::string msg_;
};
}  // namespace detail

class ReorderException : public detail::MessageException {
 public:
  explicit ReorderException(const std::string &msg) : MessageException(msg) {}
};

class TypeException : public detail::MessageException {
 public:
  explicit TypeException(const std::string &msg) : MessageException(msg) {}
  TypeException(const std::string &have, const std::string &want)
      : MessageException("Object is of type " + have + " not " + want) {}
};

class ConversionException : public detail::MessageException {
 public:
  ConversionException(const std::string &from, const std::string &to)
      : MessageException("Can not convert type " + from + " to " + to) {}
};

inline std::string ListOfKeysToString(const std::vector<std::type_index> &key) {
  std::string out = "[";
  for (size_t i = 0; i < key.size(); i++) out += (i ? ", " : "") + std::string(key[i].name());
  return out + "]";
}

template <typename KeyType>
class DirectExecutionNotAvailableException : public Exception {
 public:
  KeyType used_format_;
  std::vector<KeyType> available_formats_;
  std::string msg_;
  DirectExecutionNotAvailableException(const KeyType &used, const std::vector<KeyType> &avail)
      : used_format_(used), available_formats_(avail) {
    msg_ = "Preprocessing could not be used directly using input formats:\n " +
           ListOfKeysToString(used_format_) +
           "\nThis class can only be used with the following formats:\n ";
    for (const auto &k : available_formats_) msg_ += ListOfKeysToString(k) + "\n ";
  }
  const char *what() const noexcept override { return msg_.c_str(); }
};

class FunctionNotFoundException : public detail::MessageException {
 public:
  explicit FunctionNotFoundException(const std::string &msg) : MessageException(msg) {}
};

class NoConverterException : public detail::MessageException {
 public:
  NoConverterException()
      : MessageException("Attempting to convert a format in a preprocessing object that does "
                         "not have a Converter") {}
};

class CUDADeviceException : public detail::MessageException {
 public:
  CUDADeviceException(int available_devices, int requested_device)
      : MessageException("Attempting to use CUDA device " + std::to_string(requested_device) +
                         " when only " + std::to_string(available_devices) +
                         " CUDA devices are available\n") {}
};

class AllocationException : public detail::MessageException {
 public:
  AllocationException() : MessageException("Memory Allocation Operation Failed") {}
};

// Not in the reference (it never checks CUDA return codes): any other
// failure reported by libsb200.so.
class CUDAException : public detail::MessageException {
 public:
  CUDAException(int code, const std::string &msg)
      : MessageException("libsb200 error " + std::to_string(code) + ": " + msg), code_(code) {}
  int code() const { return code_; }

 private:
  int code_;
};

inline std::string demangle(const std::string &name) {
  int status = 0;
  char *res = abi::__cxa_demangle(name.c_str(), nullptr, nullptr, &status);
  if (status != 0 || !res) return name;
  std::string out = res;
  std::free(res);
  return out;
}
inline std::string demangle(std::type_index type) { return demangle(type.name()); }

class Identifiable {
 public:
  virtual std::type_index get_id() const = 0;
  virtual std::string get_name() const = 0;
  virtual ~Identifiable() = default;
};

template <typename IdentifiableType, typename Base>
class IdentifiableImplementation : public Base {
 public:
  using Base::Base;
  std::type_index get_id() const override { return typeid(IdentifiableType); }
  std::string get_name() const override { return utils::demangle(get_id()); }
  static std::type_index get_id_static() { return typeid(IdentifiableType); }
  static std::string get_name_static() { return utils::demangle(get_id_static()); }
};

struct Parameters {
  virtual ~Parameters() = default;
};

class Parameterizable {
 public:
  typedef Parameters ParamsType;
  virtual ~Parameterizable() = default;

 protected:
  std::unique_ptr<Parameters> params_;
};

struct TypeIndexVectorHash {
  std::size_t operator()(const std::vector<std::type_index> &v) const {
    std::size_t h = 0x9e3779b97f4a7c15ull;
    for (const auto &t : v) h = (h ^ t.hash_code()) * 0x100000001b3ull;
    return h;
  }
};

}  // namespace utils

// ======================================================================= sb200 glue
namespace sb200 {

// Translate a C-ABI return code into the exception hierarchy above.
inline void check(int rc, int device = -1) {
  if (rc == SB200_OK) return;
  const std::string msg = sb200_last_error();
  if (rc == SB200_ERR_BAD_DEVICE) {
    int cnt = 0;
    sb200_device_count(&cnt);
    throw utils::CUDADeviceException(cnt, device);
  }
  if (rc == SB200_ERR_ALLOC) throw utils::AllocationException();
  throw utils::CUDAException(rc, msg);
}

// C++ element type -> sb200_dtype code.  Value types are moved bit-exactly, so any 4/8-byte
// trivially copyable type maps to the integer code of its width; void means "no values".
template <typename T>
constexpr int dtype_of() {
  if constexpr (std::is_void_v<T>)
    return SB200_VOID;
  else if constexpr (std::is_same_v<T, float>)
    return SB200_F32;
  else if constexpr (std::is_same_v<T, double>)
    return SB200_F64;
  else if constexpr (std::is_integral_v<T> && sizeof(T) == 4)
    return std::is_signed_v<T> ? SB200_I32 : SB200_U32;
  else if constexpr (std::is_integral_v<T> && sizeof(T) == 8)
    return std::is_signed_v<T> ? SB200_I64 : SB200_U64;
  else {
    static_assert(utils::always_false<T>, "libsb200 supports 4- and 8-byte element types");
    return -1;
  }
}

template <typename T>
constexpr size_t size_of() {
  if constexpr (std::is_void_v<T>)
    return 0;
  else
    return sizeof(T);
}

// Device the host-side constructors use for their sortedness check / sort when the format
// itself lives in host memory (see format::COO / format::CSR).  Defaults to device 0.
inline int &default_device() {
  static int dev = 0;
  return dev;
}

// ---- raw device / host array helpers (all C-ABI calls) ----
template <typename T>
T *device_alloc(int device, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    void *p = nullptr;
    check(sb200_malloc(device, count * sizeof(T), &p), device);
    return static_cast<T *>(p);
  }
}
inline void device_free(int device, void *p) {
  if (p) sb200_free(device, p);
}
template <typename T>
T *upload(int device, const T *host, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!host) return nullptr;
    T *d = device_alloc<T>(device, count);
    check(sb200_memcpy_h2d(device, d, host, count * sizeof(T), nullptr), device);
    check(sb200_stream_synchronize(device, nullptr), device);
    return d;
  }
}
template <typename T>
T *download(int device, const T *dev_ptr, size_t count) {
  if constexpr (std::is_void_v<T>) {
    return nullptr;
  } else {
    if (!dev_ptr) return nullptr;
    T *h = new T[count ? count : 1];
    check(sb200_memcpy_d2h(device, h, dev_ptr, count * sizeof(T), nullptr), device);
    check(sb200_stream_synchronize(device, nullptr), device);
    return h;
  }
}

// RAII scratch allocation on a device
template <typename T>
class DeviceScratch {
 public:
  DeviceScratch(int device, size_t count) : device_(device), p_(device_alloc<T>(device, count)) {}
  DeviceScratch(int device, const T *host, size_t count)
      : device_(device), p_(upload<T>(device, host, count)) {}
  ~DeviceScratch() { device_free(device_, p_); }
  DeviceScratch(const DeviceScratch &) = delete;
  DeviceScratch &operator=(const DeviceScratch &) = delete;
  T *get() const { return p_; }
  T *release() {
    T *p = p_;
    p_ = nullptr;
    return p;
  }

 private:
  int device_;
  T *p_;
};

}  // namespace sb200

// ======================================================================= context
namespace context {

struct Context : public utils::Identifiable {
  virtual bool IsEquivalent(Context *) const = 0;
  virtual ~Context() {}
};

struct CPUContext : utils::IdentifiableImplementation<CPUContext, Context> {
  bool IsEquivalent(Context *rhs) const override {
    return rhs != nullptr && rhs->get_id() == get_id_static();
  }
};

// A CUDA device.  Like the reference, construction validates the id against the number of
// visible devices and throws utils::CUDADeviceException (cuda_context_cuda.cu:9-15); two
// CUDA contexts are equivalent iff they name the same device (:16-22).
struct CUDAContext : utils::IdentifiableImplementation<CUDAContext, Context> {
  int device_id;
  explicit CUDAContext(int did) : device_id(did) {
    int cnt = 0;
    sb200_device_count(&cnt);  // no device / no driver -> cnt == 0
    if (did < 0 || did >= cnt) throw utils::CUDADeviceException(cnt, did);
  }
  bool IsEquivalent(Context *rhs) const override {
    if (rhs == nullptr || rhs->get_id() != get_id_static()) return false;
    return static_cast<CUDAContext *>(rhs)->device_id == device_id;
  }
};

}  // namespace context

}  // namespace sparsebase
