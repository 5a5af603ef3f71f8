// Template instantiations for embedding dims 1..8 (head, loss, class sums): one translation unit
// per range keeps the per-file compile time bounded and lets the build run them in parallel.
#include "dml_head.cuh"
#include "dml_loss.cuh"
#include "dml_reduce.cuh"

namespace dml {

int head_dispatch_1_8(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s) {
  switch (D) {
    DML_HEAD_CASE(1)
    DML_HEAD_CASE(2)
    DML_HEAD_CASE(3)
    DML_HEAD_CASE(4)
    DML_HEAD_CASE(5)
    DML_HEAD_CASE(6)
    DML_HEAD_CASE(7)
    DML_HEAD_CASE(8)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int loss_dispatch_1_8(int D, int mode, int vec, bool bwd, const LossArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_LOSS_CASE(1)
    DML_LOSS_CASE(2)
    DML_LOSS_CASE(3)
    DML_LOSS_CASE(4)
    DML_LOSS_CASE(5)
    DML_LOSS_CASE(6)
    DML_LOSS_CASE(7)
    DML_LOSS_CASE(8)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int reduce_dispatch_1_8(int D, const ReduceArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_REDUCE_CASE(1)
    DML_REDUCE_CASE(2)
    DML_REDUCE_CASE(3)
    DML_REDUCE_CASE(4)
    DML_REDUCE_CASE(5)
    DML_REDUCE_CASE(6)
    DML_REDUCE_CASE(7)
    DML_REDUCE_CASE(8)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

}  // namespace dml
