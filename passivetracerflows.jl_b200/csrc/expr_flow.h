// PTF_FLOW_EXPR helper: run-time compiled velocity expressions (see expr_flow.cu).
#pragma once
#include <string>

#include "ptf_internal.h"

namespace ptf {

struct ExprFlow {
  std::string expr[3];
  bool stale = true;
  int compiled_ndim = 0;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr, kern_fill = nullptr;
  DevBuf<double> d_t;   // clock.t of the current step (TAD.jl:701,718,737: frozen for all stages)

  ExprFlow() = default;
  ExprFlow(const ExprFlow&) = delete;
  ExprFlow& operator=(const ExprFlow&) = delete;
  ~ExprFlow();
  void set(int comp, const char* e);
  bool ready(int ndim) const { return !stale && kern && compiled_ndim == ndim; }
  void compile(int ndim);
  void set_time(double t, cudaStream_t st);
  // product over a local block of nx*ny*nzl points per member whose first row / plane has global index joff / koff
  void launch(cudaStream_t st, int blocks, int nbatch, double* g0, const double* g1, const double* g2, int64_t nx,
              int64_t ny, int64_t nzl, int64_t joff, int64_t koff, const Geometry& g);
  // the fields u, v, w themselves at clock.t over the same local block (fused engines: evaluated once per step)
  void fill(cudaStream_t st, double* u, double* v, double* w, int64_t nx, int64_t ny, int64_t nzl, int64_t joff,
            int64_t koff, const Geometry& g);
};

}  // namespace ptf
