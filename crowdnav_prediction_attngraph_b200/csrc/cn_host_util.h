// Host-side error plumbing shared by the C-ABI translation units.
#pragma once
#include <stdarg.h>
#include <stdio.h>

// Records a printf-style message for cn_last_error() (thread-local) and returns 1.
int cn_set_error(const char* fmt, ...);

// Hot entry points (cn_env_step / reset, cn_policy_act, cn_gst_step / reset, cn_copy_segments) run on the handle's device and
// RESTORE the caller's current device before returning, so a host wrapper needs no device context manager around them
// (torch.cuda.device() costs ~5 us of Python per call on the train.py-contract loop).
#ifdef __CUDACC__
#include <cuda_runtime.h>
struct CnDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit CnDeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) { cudaSetDevice(device); switched = prev >= 0; }
  }
  ~CnDeviceGuard() { if (switched) cudaSetDevice(prev); }
  CnDeviceGuard(const CnDeviceGuard&) = delete;
  CnDeviceGuard& operator=(const CnDeviceGuard&) = delete;
};
#endif
