// One object file per transform length (compile with -DPTF_INST_N=256|512|1024|2048|4096): instantiates the fused
// kernels for that length and exports the launchers engine_fused.cu dispatches to.
#ifndef PTF_INST_N
#error "compile with -DPTF_INST_N=<transform length>"
#endif
#include "fused3d_kernels.cuh"
#ifdef PTF_FFT_EXPERIMENTS
namespace ptf {
#include "../../microbench/fft_experiments.cuh"
}
#endif

#define PTF_CAT2(a, b) a##b
#define PTF_CAT(a, b) PTF_CAT2(a, b)
#define FN(name) PTF_CAT(name, PTF_INST_N)

namespace ptf {

void FN(fused_prep_)() {
  prep_y<PTF_INST_N>();
  prep_x<PTF_INST_N>();
}

void FN(fused_launch_y_)(bool has_in, int fam, const void* yargs, int nb, cudaStream_t st, int n_sm) {
  launch_y<PTF_INST_N>(has_in, fam, *static_cast<const YArgs*>(yargs), nb, st, n_sm);
}

void FN(fused_launch_x_)(int vmode, const void* xargs, int nb, cudaStream_t st, int n_sm) {
  launch_x<PTF_INST_N>(vmode, *static_cast<const XArgs*>(xargs), nb, st, n_sm);
}

// ---- fused 3-D engine (engine_fused3d.cu): transform lengths up to 1024 per axis ----
#if PTF_INST_N <= 1024
void FN(fused3_prep_)() { prep3<PTF_INST_N>(); }
void FN(fused3_launch_z_)(bool has_in, int fam, const void* yargs, cudaStream_t st, int n_sm) {
  launch_z3<PTF_INST_N>(has_in, fam, *static_cast<const YArgs*>(yargs), st, n_sm);
}
void FN(fused3_launch_x_)(int vmode, const void* xargs, int nplanes, cudaStream_t st, int n_sm) {
  launch_x3<PTF_INST_N>(vmode, *static_cast<const XArgs*>(xargs), nplanes, st, n_sm);
}
void FN(fused3_launch_y_)(bool inverse, const void* y3args, cudaStream_t st, int n_sm) {
  launch_y3<PTF_INST_N>(inverse, *static_cast<const Y3Args*>(y3args), st, n_sm);
}
#endif

// transform self-test (tests/test_gpu_fused.py::test_fft_core_matches_numpy).  The timing experiments of
// profiles/r01_fft_core_experiments.md live in microbench/fft_experiments.cuh (-DPTF_FFT_EXPERIMENTS builds only).
void FN(fused_selftest_)(int dir, int count, const double2* in, double2* out, const void* twp) {
  constexpr int NN = PTF_INST_N;
  constexpr int F = 256 / Cfg<NN>::T;
  const Twiddles tw = *static_cast<const Twiddles*>(twp);
  const size_t sm = y_smem<NN>() + g_smem_pad;
  const int blocks = (count + F - 1) / F;
#ifdef PTF_FFT_EXPERIMENTS
  const size_t total = (size_t)NN * count;
  const bool timing = std::getenv("PTF_SELFTEST_TIME") != nullptr;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  const double gb = 2.0 * (double)total * 16 / 1e9;
  if (dir == 2) {
    const int pblocks = (count / 2 + F - 1) / F;
    allow_smem(k_fft_pair_test<NN>, sm);
    k_fft_pair_test<NN><<<pblocks, 256, sm>>>(in, out, count, tw);
    if (timing) {
      cudaEventRecord(e0);
      for (int r = 0; r < 10; ++r) k_fft_pair_test<NN><<<pblocks, 256, sm>>>(in, out, count, tw);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      fprintf(stderr, "[selftest_fft] pair gather+scatter via TMEM n=%d count=%d  %.4f ms  %.0f GB/s (R+W)\n", NN, count,
              ms, gb / ms * 1e3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return;
  }
#else
  if (dir == 2) throw Error(PTF_EUNSUPPORTED, "the TMEM pairing experiment needs a -DPTF_FFT_EXPERIMENTS build");
#endif
  if (dir < 0) {
    allow_smem(k_fft_test<NN, -1>, sm);
    k_fft_test<NN, -1><<<blocks, 256, sm>>>(in, out, count, tw);
  } else {
    allow_smem(k_fft_test<NN, +1>, sm);
    k_fft_test<NN, +1><<<blocks, 256, sm>>>(in, out, count, tw);
  }
#ifdef PTF_FFT_EXPERIMENTS
  if (timing && count % F == 0) {
    cudaDeviceSynchronize();
    const int reps = 10;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) {
      if (dir < 0) k_fft_test<NN, -1><<<blocks, 256, sm>>>(in, out, count, tw);
      else k_fft_test<NN, +1><<<blocks, 256, sm>>>(in, out, count, tw);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    fprintf(stderr, "[selftest_fft] n=%d count=%d  %.4f ms  %.0f GB/s (R+W)\n", NN, count, ms, gb / ms * 1e3);
    if (const char* rp = std::getenv("PTF_SELFTEST_REPEAT")) {
      const int rep = std::atoi(rp);
      allow_smem(k_fft_rate_test<NN>, sm);
      k_fft_rate_test<NN><<<blocks, 256, sm>>>(in, out, count, tw, rep);
      cudaEventRecord(e0);
      k_fft_rate_test<NN><<<blocks, 256, sm>>>(in, out, count, tw, rep);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      const double nfft = (double)count * rep;
      fprintf(stderr,
              "[selftest_fft] compute-only n=%d: %d x %d transforms in %.4f ms -> %.1f transforms/us (chip), %.0f "
              "cycles/transform/SM @1.9GHz\n",
              NN, count, rep, ms, nfft / (ms * 1e3), ms * 1e-3 * 1.9e9 * 148 / nfft);
    }
    for (int pair = 0; pair < 2; ++pair) {
      if (pair == 0) allow_smem(k_fft_gather_test<NN, 0>, sm);
      else allow_smem(k_fft_gather_test<NN, 1>, sm);
      cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) {
        if (pair == 0) k_fft_gather_test<NN, 0><<<blocks, 256, sm>>>(in, out, count, tw);
        else k_fft_gather_test<NN, 1><<<blocks, 256, sm>>>(in, out, count, tw);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= reps;
      fprintf(stderr, "[selftest_fft] gather-in (pair=%d)       %.4f ms  %.0f GB/s (R+W)\n", pair, ms, gb / ms * 1e3);
    }
    // the experiments overwrite `out`: recompute the requested transform last
    if (dir < 0) k_fft_test<NN, -1><<<blocks, 256, sm>>>(in, out, count, tw);
    else k_fft_test<NN, +1><<<blocks, 256, sm>>>(in, out, count, tw);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
#endif
}

}  // namespace ptf
