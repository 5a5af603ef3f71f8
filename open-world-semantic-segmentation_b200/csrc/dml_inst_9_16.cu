// Template instantiations for embedding dims 9..16 (head, loss, class sums): one translation unit
// per range keeps the per-file compile time bounded and lets the build run them in parallel.
#include "dml_head.cuh"
#include "dml_loss.cuh"
#include "dml_reduce.cuh"

namespace dml {

int head_dispatch_9_16(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s) {
  switch (D) {
    DML_HEAD_CASE(9)
    DML_HEAD_CASE(10)
    DML_HEAD_CASE(11)
    DML_HEAD_CASE(12)
    DML_HEAD_CASE(13)
    DML_HEAD_CASE(14)
    DML_HEAD_CASE(15)
    DML_HEAD_CASE(16)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int loss_dispatch_9_16(int D, int mode, int vec, bool bwd, const LossArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_LOSS_CASE(9)
    DML_LOSS_CASE(10)
    DML_LOSS_CASE(11)
    DML_LOSS_CASE(12)
    DML_LOSS_CASE(13)
    DML_LOSS_CASE(14)
    DML_LOSS_CASE(15)
    DML_LOSS_CASE(16)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int reduce_dispatch_9_16(int D, const ReduceArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_REDUCE_CASE(9)
    DML_REDUCE_CASE(10)
    DML_REDUCE_CASE(11)
    DML_REDUCE_CASE(12)
    DML_REDUCE_CASE(13)
    DML_REDUCE_CASE(14)
    DML_REDUCE_CASE(15)
    DML_REDUCE_CASE(16)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

}  // namespace dml
