/* allreduce_ref.c — ORACLE (test infrastructure): the bucket arithmetic of the reference stack in plain C.
 *
 * Restates, for fp32 buckets, what the reference's dependency computes per gradient bucket
 * (torch c10d):
 *   oracle_flatten_scale   Reducer::mark_variable_ready_dense, reducer.hpp:276 (copy fused with *1/ws,
 *                          default_comm_hooks.hpp:36-41)
 *   oracle_allreduce_sum   ProcessGroup::allreduce(SUM), ProcessGroupNCCL.hpp:856 / ProcessGroupGloo.hpp:336,
 *                          with OUR fixed summation order (rank 0..ws-1, fp32) so that the CUDA kernels can be
 *                          compared bit for bit
 *   oracle_unflatten       Reducer::copy_bucket_to_grad, reducer.hpp:330
 * Used by tests/ for full-size parity and by bench.py's cpu_baseline leg (kind "port").  Never linked into
 * or called from the product library.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp, -ffp-contract=off so no FMA changes the rounding).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bucket[off[i] + j] = grads[i][j] * scale */
void oracle_flatten_scale(const float* const* grads, const uint64_t* off, const uint64_t* len, int n, float scale,
                          float* bucket) {
  for (int i = 0; i < n; ++i) {
    const float* g = grads[i];
    float* b = bucket + off[i];
    const int64_t m = (int64_t)len[i];
#pragma omp parallel for schedule(static) if (m > 65536)
    for (int64_t j = 0; j < m; ++j) b[j] = g[j] * scale;
  }
}

/* out[j] = ((in[0][j] + in[1][j]) + in[2][j]) + ...   (fp32, rank order) */
void oracle_allreduce_sum(const float* const* in, int ws, uint64_t n, float* out) {
#pragma omp parallel for schedule(static) if (n > 65536)
  for (int64_t j = 0; j < (int64_t)n; ++j) {
    float acc = in[0][j];
    for (int r = 1; r < ws; ++r) acc = acc + in[r][j];
    out[j] = acc;
  }
}

/* grads[i][j] = bucket[off[i] + j] */
void oracle_unflatten(const float* bucket, const uint64_t* off, const uint64_t* len, int n, float* const* grads) {
  for (int i = 0; i < n; ++i) memcpy(grads[i], bucket + off[i], (size_t)len[i] * sizeof(float));
}

/* One whole step for `ws` ranks living in one address space: every rank's grads are flattened+scaled,
 * summed, and written back to every rank.  scratch: ws * total floats.  Returns bytes of gradient
 * processed per rank (for throughput accounting). */
uint64_t oracle_ddp_step(float* const* const* grads /* [ws][n] */, const uint64_t* off, const uint64_t* len, int n,
                         int ws, uint64_t total, float* scratch) {
  const float scale = 1.0f / (float)ws;
  for (int r = 0; r < ws; ++r)
    oracle_flatten_scale((const float* const*)grads[r], off, len, n, scale, scratch + (uint64_t)r * total);
  const float* in[64] = {0};
  for (int r = 0; r < ws; ++r) in[r] = scratch + (uint64_t)r * total;
  oracle_allreduce_sum(in, ws, total, scratch); /* in-place into rank 0's flat buffer: safe, element-wise */
  for (int r = 0; r < ws; ++r) oracle_unflatten(scratch, off, len, n, grads[r]);
  uint64_t bytes = 0;
  for (int i = 0; i < n; ++i) bytes += len[i] * sizeof(float);
  return bytes;
}
