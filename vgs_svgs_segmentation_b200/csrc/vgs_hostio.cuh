// vgs_hostio.cuh — host side of the boundary: moving LARGE PAGEABLE host buffers (the pcl::PointCloud the drop-in classes
// are handed, the std::vector results they return) to and from the device.
//
// A plain cudaMemcpy of pageable memory is one driver thread copying through a small pinned bounce buffer: ~10 GB/s, and the
// first touch of a fresh destination (page faults) is paid by that one thread as well.  Here a ring of pinned staging chunks
// is filled / drained by a few host threads, each with its own copy stream: the host-side memcpys (and the page faults of
// a fresh std::vector) run in parallel and overlap the DMA.  Pinned callers (bench.py's e2e, the slab group) never come
// here: their buffers go straight to cudaMemcpyAsync.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace vgs_hostio {

constexpr int kThreads = 4;                 // host threads per transfer (each ~8-10 GB/s of memcpy)
constexpr size_t kChunk = 4u << 20;         // staging chunk
constexpr size_t kMinBytes = 8u << 20;      // below this a plain copy wins (thread start-up)

struct Engine {
  std::mutex mu;                            // one staged transfer at a time per process
  bool ready = false, broken = false;
  int device = -1;
  void* pin[kThreads][2] = {};
  cudaStream_t st[kThreads] = {};
  cudaEvent_t ev[kThreads][2] = {};
  cudaEvent_t ev_begin = nullptr, ev_end[kThreads] = {};
};
inline Engine& engine() { static Engine e; return e; }

// true when p is ordinary (unregistered) host memory
inline bool pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

inline bool prepare(Engine& e, int device) {
  static const bool off = [] { const char* v = std::getenv("VGS_B200_NO_STAGED_COPY"); return v && v[0] == '1'; }();   // A/B knob
  if (e.broken || off) return false;
  if (e.ready && e.device == device) return true;
  if (e.ready) {   // streams and events belong to a device: rebuild them (the pinned chunks are portable)
    for (int t = 0; t < kThreads; t++) {
      cudaStreamSynchronize(e.st[t]);        // a chunk may still be on the bus
      cudaStreamDestroy(e.st[t]); cudaEventDestroy(e.ev_end[t]);
      for (int s = 0; s < 2; s++) cudaEventDestroy(e.ev[t][s]);
    }
    cudaEventDestroy(e.ev_begin);
    e.ready = false;
  }
  bool ok = true;
  for (int t = 0; t < kThreads && ok; t++) {
    for (int s = 0; s < 2 && ok; s++) {
      if (!e.pin[t][s]) ok = cudaHostAlloc(&e.pin[t][s], kChunk, cudaHostAllocPortable) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&e.ev[t][s], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaStreamCreateWithFlags(&e.st[t], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&e.ev_end[t], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaEventCreateWithFlags(&e.ev_begin, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { cudaGetLastError(); e.broken = true; return false; }
  e.ready = true; e.device = device;
  return true;
}

// host (pageable) -> device, ordered after the work already on `stream`; on return the caller's buffer has been read
// completely and `stream` waits for the last chunk.  Falls back to cudaMemcpyAsync for small / pinned buffers.
inline cudaError_t to_device(void* dst, const void* src, size_t bytes, int device, cudaStream_t stream) {
  Engine& e = engine();
  if (bytes < kMinBytes || !pageable(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  std::lock_guard<std::mutex> lock(e.mu);
  if (!prepare(e, device)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  cudaError_t r = cudaEventRecord(e.ev_begin, stream);
  if (r != cudaSuccess) return r;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  std::atomic<int> err{0};
  auto worker = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamWaitEvent(e.st[t], e.ev_begin, 0) != cudaSuccess) { err = 1; return; }
    int it = 0;
    for (size_t c = (size_t)t; c < nchunks; c += kThreads, it++) {
      const int s = it & 1;
      const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      if (cudaEventSynchronize(e.ev[t][s]) != cudaSuccess) { err = 1; return; }     // the chunk's previous DMA has drained
      std::memcpy(e.pin[t][s], (const char*)src + off, len);
      if (cudaMemcpyAsync((char*)dst + off, e.pin[t][s], len, cudaMemcpyHostToDevice, e.st[t]) != cudaSuccess ||
          cudaEventRecord(e.ev[t][s], e.st[t]) != cudaSuccess) { err = 1; return; }
    }
    if (cudaEventRecord(e.ev_end[t], e.st[t]) != cudaSuccess) err = 1;
  };
  std::thread th[kThreads];
  for (int t = 1; t < kThreads; t++) {
    try { th[t] = std::thread(worker, t); }
    catch (...) { worker(t); }          // no thread to be had (no exception may cross the C ABI): this share runs inline
  }
  worker(0);
  for (int t = 1; t < kThreads; t++) if (th[t].joinable()) th[t].join();
  if (err) { cudaGetLastError(); return cudaErrorUnknown; }
  for (int t = 0; t < kThreads; t++)
    if ((r = cudaStreamWaitEvent(stream, e.ev_end[t], 0)) != cudaSuccess) return r;
  return cudaSuccess;
}

// device -> host (pageable), ordered after the work already on `stream`; SYNCHRONOUS: the caller's buffer is complete on
// return.  Falls back to cudaMemcpyAsync + synchronize.
inline cudaError_t to_host(void* dst, const void* src, size_t bytes, int device, cudaStream_t stream) {
  Engine& e = engine();
  auto plain = [&]() -> cudaError_t {
    cudaError_t r = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream);
    return r != cudaSuccess ? r : cudaStreamSynchronize(stream);
  };
  if (bytes < kMinBytes || !pageable(dst)) return plain();
  std::lock_guard<std::mutex> lock(e.mu);
  if (!prepare(e, device)) return plain();
  cudaError_t r = cudaEventRecord(e.ev_begin, stream);
  if (r != cudaSuccess) return r;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  std::atomic<int> err{0};
  auto worker = [&](int t) {
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamWaitEvent(e.st[t], e.ev_begin, 0) != cudaSuccess) { err = 1; return; }
    size_t prev_off = 0, prev_len = 0;
    int it = 0, prev_s = -1;
    for (size_t c = (size_t)t; c < nchunks; c += kThreads, it++) {
      const int s = it & 1;
      const size_t off = c * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      if (cudaMemcpyAsync(e.pin[t][s], (const char*)src + off, len, cudaMemcpyDeviceToHost, e.st[t]) != cudaSuccess ||
          cudaEventRecord(e.ev[t][s], e.st[t]) != cudaSuccess) { err = 1; return; }
      if (prev_s >= 0) {     // drain the previous chunk while this one is on the bus
        if (cudaEventSynchronize(e.ev[t][prev_s]) != cudaSuccess) { err = 1; return; }
        std::memcpy((char*)dst + prev_off, e.pin[t][prev_s], prev_len);
      }
      prev_s = s; prev_off = off; prev_len = len;
    }
    if (prev_s >= 0) {
      if (cudaEventSynchronize(e.ev[t][prev_s]) != cudaSuccess) { err = 1; return; }
      std::memcpy((char*)dst + prev_off, e.pin[t][prev_s], prev_len);
    }
  };
  std::thread th[kThreads];
  for (int t = 1; t < kThreads; t++) {
    try { th[t] = std::thread(worker, t); }
    catch (...) { worker(t); }          // no thread to be had (no exception may cross the C ABI): this share runs inline
  }
  worker(0);
  for (int t = 1; t < kThreads; t++) if (th[t].joinable()) th[t].join();
  if (err) { cudaGetLastError(); return cudaErrorUnknown; }
  return cudaSuccess;
}

}  // namespace vgs_hostio
