// Error reporting and launch accounting shared by all entry points.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace clslam {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CLSLAM_PDL");
    v = (e && e[0] == '1') ? 1 : 0;      // measured: no gain inside the CUDA graph (profiles/), off by default
  }
  return v == 1;
}

int finish_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return -100 - (int)e;
  }
  return 0;
}

}  // namespace clslam

extern "C" {
const char* clslam_last_error(void) { return clslam::g_err; }
int clslam_abi_version(void) { return CLSLAM_ABI_VERSION; }
long long clslam_launch_count(void) { return clslam::g_launches.load(); }
// ---- timing events ---------------------------------------------------------------------------------
void* clslam_event_create(void) {
  cudaEvent_t ev = nullptr;
  const cudaError_t e = cudaEventCreate(&ev);
  if (e != cudaSuccess) { clslam::set_error("event_create: %s", cudaGetErrorString(e)); return nullptr; }
  return ev;
}
int clslam_event_destroy(void* ev) {
  if (ev) cudaEventDestroy(reinterpret_cast<cudaEvent_t>(ev));
  return 0;
}
int clslam_event_record(void* ev, void* stream) {
  cudaStream_t st = clslam::as_stream(stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  // inside a capture the record becomes an event-record NODE of the graph (external event): every replay stamps it
  const cudaError_t e = cudaEventRecordWithFlags(reinterpret_cast<cudaEvent_t>(ev), st,
                                                 cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
  if (e != cudaSuccess) { clslam::set_error("event_record: %s", cudaGetErrorString(e)); return -100 - (int)e; }
  return 0;
}
int clslam_event_elapsed_ms(void* ev0, void* ev1, float* ms) {
  const cudaError_t e = cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(ev0), reinterpret_cast<cudaEvent_t>(ev1));
  if (e != cudaSuccess) { clslam::set_error("event_elapsed: %s", cudaGetErrorString(e)); return -100 - (int)e; }
  return 0;
}
int clslam_set_device(int device) {
  const cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { clslam::set_error("set_device(%d): %s", device, cudaGetErrorString(e)); return -100 - (int)e; }
  return 0;
}
}
