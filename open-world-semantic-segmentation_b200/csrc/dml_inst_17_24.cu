// Template instantiations for embedding dims 17..24 (head, loss, class sums): one translation unit
// per range keeps the per-file compile time bounded and lets the build run them in parallel.
#include "dml_head.cuh"
#include "dml_loss.cuh"
#include "dml_reduce.cuh"

namespace dml {

int head_dispatch_17_24(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s) {
  switch (D) {
    DML_HEAD_CASE(17)
    DML_HEAD_CASE(18)
    DML_HEAD_CASE(19)
    DML_HEAD_CASE(20)
    DML_HEAD_CASE(21)
    DML_HEAD_CASE(22)
    DML_HEAD_CASE(23)
    DML_HEAD_CASE(24)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int loss_dispatch_17_24(int D, int mode, int vec, bool bwd, const LossArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_LOSS_CASE(17)
    DML_LOSS_CASE(18)
    DML_LOSS_CASE(19)
    DML_LOSS_CASE(20)
    DML_LOSS_CASE(21)
    DML_LOSS_CASE(22)
    DML_LOSS_CASE(23)
    DML_LOSS_CASE(24)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int reduce_dispatch_17_24(int D, const ReduceArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_REDUCE_CASE(17)
    DML_REDUCE_CASE(18)
    DML_REDUCE_CASE(19)
    DML_REDUCE_CASE(20)
    DML_REDUCE_CASE(21)
    DML_REDUCE_CASE(22)
    DML_REDUCE_CASE(23)
    DML_REDUCE_CASE(24)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

}  // namespace dml
