#pragma once
// Timing experiments on the hand-written FFT core (recorded in profiles/r01_fft_core_experiments.md): compute-only rate,
// strided / paired gathers, the TMEM pairing unit test.  NOT part of the product library: compiled into
// ptf_selftest_fft only when the library is built with -DPTF_FFT_EXPERIMENTS
// (`python passivetracerflows.jl_b200/build.py --variant exp PTF_FFT_EXPERIMENTS=1`, then PTF_LIB_PATH=...libptf_b200_exp.so
// for tools/fft_bw.py, tools/pair_check.py, tools/sanitize_run.py).  Included by csrc/fused_inst.cu inside namespace ptf.
namespace {

// experiment: compute-only rate of the transform (REP transforms per load/store)
template <int N>
__global__ void __launch_bounds__(256, 2) k_fft_rate_test(const double2* __restrict__ in, double2* __restrict__ out,
                                                          int count, Twiddles tw, int rep) {
  constexpr int T = Cfg<N>::T, F = 256 / T, PADN = Cfg<N>::PADN;
  extern __shared__ double2 smem[];
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int id = blockIdx.x * F + grp;
  const size_t base = (size_t)id * N;
  double2 v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = in[base + t + T * e];
#pragma unroll 1
  for (int r = 0; r < rep; ++r) {
    fft::fft_cta<N, -1>(v, smem + grp * PADN, t, tw);
    double2 w[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) w[e] = v[out_slot<N>(e)];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = make_double2(w[e].x * 0.015625, w[e].y * 0.015625);
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) out[base + t + T * e] = v[e];
}

// experiment: same transform, but the input is read as column `id` of a row-major [N][count] matrix (16-byte gathers,
// the access pattern of k_fused_y's P^x read / k_fused_x's A,B read); PAIR = 1 reads 32-byte lane pairs
template <int N, int PAIR>
__global__ void __launch_bounds__(256, 2) k_fft_gather_test(const double2* __restrict__ in, double2* __restrict__ out,
                                                            int count, Twiddles tw) {
  constexpr int T = Cfg<N>::T, F = 256 / T, PADN = Cfg<N>::PADN;
  extern __shared__ double2 smem[];
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int id = blockIdx.x * F + grp;
  const size_t base = (size_t)id * N;
  double2 v[16];
  if (PAIR == 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __ldcg(in + (size_t)(t + T * e) * count + id);
  } else {
    // lanes (2j, 2j+1) read the two halves of one 32-byte sector: element index t>>1 + ..., column 2*id' + (t&1)
    const int tt = t >> 1, c = t & 1;
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __ldcg(in + (size_t)(tt + (T / 2) * e) * count + 2 * (id / 2) + c);
  }
  fft::fft_cta<N, -1>(v, smem + grp * PADN, t, tw);
#pragma unroll
  for (int e = 0; e < 16; ++e) __stcg(out + base + t + T * e, v[out_slot<N>(e)]);
}

// experiment / unit test of the pairing machinery: columns (2j, 2j+1) of a row-major [N][count] matrix are gathered
// with 256-bit loads, the odd column is parked in TMEM while the even one is transformed, and the two results are
// written back as 32-byte pairs (out[N][count], transform along axis 0).
template <int N>
__global__ void __launch_bounds__(256, 2) k_fft_pair_test(const double2* __restrict__ in, double2* __restrict__ out,
                                                          int count, Twiddles tw) {
  constexpr int T = Cfg<N>::T, F = 256 / T, PADN = Cfg<N>::PADN;
  extern __shared__ double2 smem[];
  __shared__ uint32_t tslot;
  const uint32_t tbase = tmem::alloc_cta<256>(&tslot);
  const uint32_t ta = tmem::warp_addr(tbase, 128);
  const int grp = threadIdx.x / T, t = threadIdx.x % T;
  const int pair = blockIdx.x * F + grp;
  const bool active = 2 * pair < count;
  const size_t c0 = 2 * (size_t)(active ? pair : 0);
  double2 v[16], w[16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // two batches of 8 pairs: 64 registers in flight, odd column parked at once
    double2 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tmem::ldg256(in + (size_t)(t + T * (8 * h + j)) * count + c0, v[8 * h + j], o[j]);
    tmem::st4(ta + 32 * h, o[0], o[1], o[2], o[3]);
    tmem::st4(ta + 32 * h + 16, o[4], o[5], o[6], o[7]);
  }
  tmem::wait_st();
  fft::fft_cta<N, -1>(v, smem + grp * PADN, t, tw);
  tmem::park16(ta + 64, v, [](int i) { return out_slot<N>(i); });   // result of column 0, natural order
  tmem::fetch16(ta, w);
  fft::fft_cta<N, -1>(w, smem + grp * PADN, t, tw);
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // fetch 4 parked results at a time: never more than 16 extra registers live
    double2 a0, a1, a2, a3;
    tmem::ld4(ta + 64 + 16 * q, a0, a1, a2, a3);
    if (active) {
      tmem::stg256(out + (size_t)(t + T * (4 * q + 0)) * count + c0, a0, w[out_slot<N>(4 * q + 0)]);
      tmem::stg256(out + (size_t)(t + T * (4 * q + 1)) * count + c0, a1, w[out_slot<N>(4 * q + 1)]);
      tmem::stg256(out + (size_t)(t + T * (4 * q + 2)) * count + c0, a2, w[out_slot<N>(4 * q + 2)]);
      tmem::stg256(out + (size_t)(t + T * (4 * q + 3)) * count + c0, a3, w[out_slot<N>(4 * q + 3)]);
    }
  }
  tmem::free_cta<256>(tbase);
}


}  // namespace
