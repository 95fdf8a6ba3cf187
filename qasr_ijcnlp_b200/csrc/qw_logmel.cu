// placeholder, replaced below
#include "../../include/qw.h"
#include "qw_common.cuh"
extern "C" {
size_t qw_log_mel_workspace_bytes(int B, int n_samples, int n_mels) { return 0; }
int qw_log_mel(const float*, const float*, float*, void*, size_t, int, int, int, void*) {
  qw::set_error("qw_log_mel: not built yet");
  return -2;
}
}
