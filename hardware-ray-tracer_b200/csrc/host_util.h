// Host-side helpers: error propagation without exceptions across the C ABI, grow-only device buffers.
#pragma once
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "compat.cuh"

namespace brt {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
struct LimitError : std::runtime_error {
  explicit LimitError(const std::string& m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    throw CudaError(std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
  }
}
#define BRT_CUDA(x) ::brt::cuda_check((x), #x, __FILE__, __LINE__)
#define BRT_CHECK_LAUNCH() ::brt::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// device allocation that only ever grows (scratch reused across per-frame rebuilds)
class DevBuf {
 public:
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p_(o.p_), cap_(o.cap_) { o.p_ = nullptr; o.cap_ = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p_ = o.p_; cap_ = o.cap_; o.p_ = nullptr; o.cap_ = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p_) cudaFree(p_);
    p_ = nullptr;
    cap_ = 0;
  }
  // contents are NOT preserved when the buffer grows
  void ensure(size_t bytes) {
    if (bytes <= cap_) return;
    release();
    size_t want = bytes + bytes / 8 + 256;
    BRT_CUDA(cudaMalloc(&p_, want));
    cap_ = want;
  }
  template <class T> T* as() const { return static_cast<T*>(p_); }
  void* ptr() const { return p_; }
  size_t capacity() const { return cap_; }

 private:
  void* p_ = nullptr;
  size_t cap_ = 0;
};

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace brt
