// Shared internals of libfo_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fo_b200.h"

namespace fo {

void set_error(const char* fmt, ...);
void count_launch(uint64_t n = 1);

#define FO_CUDA_TRY(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FO_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// ---- packed agent table (written by fo_agents_pack, read by the metric kernels) ----------------
// [ float4 s0[A*Tp] | float4 s1[A*Tp] | AgentParams prm[A] ]
//   s0 = (px, py, cos yaw, sin yaw)          s1 = (yaw, v, 1/(sqrt2 sigma_x), 1/(sqrt2 sigma_y))
struct __align__(16) AgentParams {
  int32_t n_states;
  int32_t model;   // 0 = unprotected (LR1S ego / pedestrian logit), 1 = protected (LR4S both), 2 = none (harm 1)
  float hl, hw;    // half extents of the unbuffered agent.shape  (DCE, BE)
  float hlb;       // half of the buffered prediction length     (CP front/back points)
  float ke, ko;    // m_o/(m_e+m_o), m_e/(m_e+m_o)               (harm_model.py:96-97)
  float pad;
};

struct AgentTableView {
  const float4* s0;
  const float4* s1;
  const AgentParams* prm;
};

__host__ __device__ inline size_t agent_table_bytes(int A, int Tp) {
  return (size_t)A * Tp * 2 * sizeof(float4) + (size_t)A * sizeof(AgentParams);
}
__host__ __device__ inline AgentTableView agent_table_view(const void* base, int A, int Tp) {
  AgentTableView v;
  v.s0 = reinterpret_cast<const float4*>(base);
  v.s1 = v.s0 + (size_t)A * Tp;
  v.prm = reinterpret_cast<const AgentParams*>(v.s1 + (size_t)A * Tp);
  return v;
}

}  // namespace fo
