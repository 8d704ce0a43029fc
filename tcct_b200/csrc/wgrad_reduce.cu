// One launch folds the partial sums of EVERY tcgen05 weight-gradient kernel of a backward pass into the gradients
// (tcct_wgrad_reduce_batch): the jobs travel by value in the kernel parameters, a block finds its job by binary search over the
// block offsets.  See csrc/wgrad_reduce.cuh.
#include "wgrad_reduce.cuh"

#define RJ_MAX 128

struct ReduceJob {            // == tcct_reduce_job of include/tcct_b200.h
  const float* ws;
  float* dw;
  int kind;                   // 0: conv (p0 = S, p1 = KA, p2 = KL); 1: 1x1 / linear (p0 = N, p1 = K, p2 = row stride of dW)
  int nparts;
  int p0, p1, p2;
  int reserved;
};
struct ReduceBatch {
  ReduceJob job[RJ_MAX];
  int start[RJ_MAX + 1];      // first block of every job
  int n;
};

__global__ void __launch_bounds__(256) wgrad_reduce_batch_kernel(const __grid_constant__ ReduceBatch rb) {
  __shared__ float s_part[8 * 32];
  int lo = 0, hi = rb.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (rb.start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ReduceJob& j = rb.job[lo];
  const int block = blockIdx.x - rb.start[lo], nblocks = rb.start[lo + 1] - rb.start[lo];
  if (j.kind == 0) wgrad_line_reduce_body(j.ws, j.nparts, j.p0, j.p1, j.p2, j.dw, block);
  else wgrad_gemm_reduce_body<8>(j.ws, j.nparts, j.p0, j.p1, j.p2, j.dw, block, nblocks, s_part);
}

extern "C" int tcct_reduce_job_size(void) { return (int)sizeof(ReduceJob); }

// jobs: host array of n tcct_reduce_job records (copied into the launch parameters); launches ceil(n / 128) kernels
extern "C" int tcct_wgrad_reduce_batch(const void* jobs_host, int n, void* stream) {
  TCCT_CHECK_ARG(n >= 0 && (n == 0 || jobs_host != nullptr), "wgrad_reduce_batch: bad job list");
  const ReduceJob* jobs = (const ReduceJob*)jobs_host;
  const int sms = tcct_num_sms();
  for (int first = 0; first < n; first += RJ_MAX) {
    ReduceBatch rb;
    rb.n = n - first < RJ_MAX ? n - first : RJ_MAX;
    int blocks = 0;
    for (int i = 0; i < rb.n; i++) {
      rb.job[i] = jobs[first + i];
      TCCT_CHECK_ARG(rb.job[i].ws && rb.job[i].dw && rb.job[i].nparts > 0 && (rb.job[i].kind == 0 || rb.job[i].kind == 1), "wgrad_reduce_batch: bad job %d", first + i);
      rb.start[i] = blocks;
      blocks += rb.job[i].kind == 0 ? wgrad_line_reduce_blocks(rb.job[i].p0) : wgrad_gemm_reduce_blocks(rb.job[i].p0, rb.job[i].p1, sms);
    }
    rb.start[rb.n] = blocks;
    wgrad_reduce_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rb);
    tcct_count_launch();
  }
  TCCT_CHECK_LAUNCH("wgrad_reduce_batch");
  return TCCT_OK;
}
