// Template instantiations for embedding dims 25..32 (head, loss, class sums): one translation unit
// per range keeps the per-file compile time bounded and lets the build run them in parallel.
#include "dml_head.cuh"
#include "dml_loss.cuh"
#include "dml_reduce.cuh"

namespace dml {

int head_dispatch_25_32(int D, int mode, int vec, bool extra, const HeadArgs& a, cudaStream_t s) {
  switch (D) {
    DML_HEAD_CASE(25)
    DML_HEAD_CASE(26)
    DML_HEAD_CASE(27)
    DML_HEAD_CASE(28)
    DML_HEAD_CASE(29)
    DML_HEAD_CASE(30)
    DML_HEAD_CASE(31)
    DML_HEAD_CASE(32)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int loss_dispatch_25_32(int D, int mode, int vec, bool bwd, const LossArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_LOSS_CASE(25)
    DML_LOSS_CASE(26)
    DML_LOSS_CASE(27)
    DML_LOSS_CASE(28)
    DML_LOSS_CASE(29)
    DML_LOSS_CASE(30)
    DML_LOSS_CASE(31)
    DML_LOSS_CASE(32)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

int reduce_dispatch_25_32(int D, const ReduceArgs& a, int gx, cudaStream_t s) {
  switch (D) {
    DML_REDUCE_CASE(25)
    DML_REDUCE_CASE(26)
    DML_REDUCE_CASE(27)
    DML_REDUCE_CASE(28)
    DML_REDUCE_CASE(29)
    DML_REDUCE_CASE(30)
    DML_REDUCE_CASE(31)
    DML_REDUCE_CASE(32)
    default:
      return DML_ERR_UNSUPPORTED_DIM;
  }
}

}  // namespace dml
