// kc_api.cu — process-wide plumbing of the C-ABI: error text, device selection, accelerators.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "kc_common.cuh"

namespace kc {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int32_t cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  cudaGetLastError();  // clear the sticky-less error state
  set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file,
            line, what);
  return e == cudaErrorMemoryAllocation ? KC_ERR_OOM : KC_ERR_CUDA;
}

static std::once_flag g_once;
static int g_dev_rc = KC_ERR_CUDA;
static int g_sms = 148;
static int g_dev = 0;  // the device every entry point runs on (chosen once per process)
static std::string g_dev_err;

// One process per GPU: under torchrun the rank's device is LOCAL_RANK (KOMPASS_B200_DEVICE
// overrides). There is no CPU fallback: no usable device => every create() fails loudly.
int32_t ensure_device() {
  std::call_once(g_once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      cudaGetLastError();
      g_dev_err = std::string("no usable CUDA device: ") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                  " (kompass_b200 has no CPU fallback)";
      return;
    }
    int dev = 0;
    const char *env = getenv("KOMPASS_B200_DEVICE");
    if (!env) env = getenv("LOCAL_RANK");
    if (env) dev = atoi(env) % n;
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) {
      g_dev_err = std::string("cudaSetDevice failed: ") + cudaGetErrorString(e);
      return;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) g_sms = prop.multiProcessorCount;
    g_dev = dev;
    g_dev_rc = KC_OK;
  });
  if (g_dev_rc != KC_OK) {
    set_error("%s", g_dev_err.c_str());
    return g_dev_rc;
  }
  // The current device is per-thread state of the runtime API: a thread that did not run the
  // call_once above (a ROS executor thread, a Python worker) starts on device 0. Every C-ABI entry
  // point comes through here, so re-assert the process's device on the calling thread (a no-op
  // inside the runtime when it is already current).
  static thread_local bool t_set = false;
  if (!t_set) {
    cudaError_t e = cudaSetDevice(g_dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    t_set = true;
  }
  return KC_OK;
}

int device_index() { return g_dev; }

int sm_count() { return g_sms; }

}  // namespace kc

// FP32 FMA micro-benchmark: the measured CUDA-core roofline denominator (MEASURED_PEAKS.json only
// carries HBM and BF16 figures). 16 independent FFMA chains per thread, explicit intrinsics because
// the library is built with -fmad=false.
__global__ void k_fp32_peak(float *out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
  const float m = 1.000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

extern "C" {

// returns measured dense FP32 throughput in TFLOP/s (FMA = 2 FLOP), best of 5 launches
int32_t kc_debug_fp32_peak_tflops(float *tflops) {
  KC_REQUIRE(tflops, KC_ERR_INVALID_ARG, "null output");
  KC_TRY(kc::ensure_device());
  struct Scope {  // released on every return path
    float *d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scope() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (d) cudaFree(d);
    }
  } s;
  KC_CUDA(cudaMalloc(&s.d, 4));
  KC_CUDA(cudaEventCreate(&s.e0));
  KC_CUDA(cudaEventCreate(&s.e1));
  const int iters = 8192, threads = 256, blocks = kc::sm_count() * 16;
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    KC_CUDA(cudaEventRecord(s.e0));
    k_fp32_peak<<<blocks, threads>>>(s.d, iters);
    KC_CUDA(cudaGetLastError());
    KC_CUDA(cudaEventRecord(s.e1));
    KC_CUDA(cudaEventSynchronize(s.e1));
    float ms = 0.0f;
    KC_CUDA(cudaEventElapsedTime(&ms, s.e0, s.e1));
    if (rep > 0 && ms > 0.0f && ms < best) best = ms;
  }
  KC_REQUIRE(best < 1e29f, KC_ERR_CUDA, "FP32 peak micro-benchmark produced no valid timing");
  const double flop = (double)blocks * threads * (double)iters * 16.0 * 2.0;
  *tflops = (float)(flop / (best * 1e-3) / 1e12);
  return KC_OK;
}

const char *kc_last_error(void) { return kc::g_err.c_str(); }

const char *kc_version(void) { return "kompass_b200 0.1.0 sm_100a"; }

int32_t kc_available_accelerators(char *buf, int32_t buf_len) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  std::string s;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess) {
      s += prop.name;
      s += "\n";
    }
  }
  if (buf && buf_len > 0) {
    strncpy(buf, s.c_str(), (size_t)buf_len - 1);
    buf[buf_len - 1] = '\0';
  }
  return n;
}

}  // extern "C"
