// Expected counts and the background-marginalised Poisson likelihood.
//
// Replaces xpsi/tools/compute_expected_counts.pyx:66-197 and
// xpsi/likelihoods/default_background_marginalisation.pyx:84-447,450-734.
//
// One warp per (parameter vector, channel):
//   * the channel's pulse of each component is splined in phase (the global
//     phase interpolant), lane j integrates data phase bins j, j+32, ... (BPL per lane) with the
//     reference's shift / wrap / clip rules;
//   * the Newton search for the ML background uses warp reductions over bins;
//   * the marginal integral replaces GSL CQUAD by a fixed-order rule that is
//     converged far below the reference's epsrel: the log-integrand is concave
//     in B, so the warp first brackets the super-level set {x(B) >= max-45} by
//     probing 32 points at a time (zooming while it is under-resolved) and then
//     applies 12 panels x 8-point Gauss-Legendre with one node per lane.
#include "common.cuh"
#include "kernels.h"

namespace xb {

constexpr int kWarpsPerBlock = 4;
#ifndef XB_MARG_PANELS
#define XB_MARG_PANELS 12
#endif
constexpr int kPanels = XB_MARG_PANELS;      // 8-point Gauss-Legendre panels over the bracketed set (a multiple of 4)
constexpr int kMaxBPL = 4;                 // bins per lane: up to 128 data phase bins

__constant__ double c_gl8_x[8] = {
    -9.60289856497536176e-01, -7.96666477413626728e-01, -5.25532409916328991e-01,
    -1.83434642495649780e-01, 1.83434642495649780e-01,  5.25532409916328991e-01,
    7.96666477413626728e-01,  9.60289856497536176e-01};
__constant__ double c_gl8_w[8] = {
    1.01228536290377064e-01, 2.22381034453374427e-01, 3.13706645877886880e-01,
    3.62683783378361657e-01, 3.62683783378361657e-01, 3.13706645877886880e-01,
    2.22381034453374427e-01, 1.01228536290377064e-01};

// exact integral of the phase spline between a <= b (both inside the node range)
__device__ __forceinline__ double spline_integ(const double* x, const double* c, int n, double a,
                                               double b) {
  if (!(b > a)) return 0.0;
  const int ia = interval_search(x, n, a);
  const int ib = interval_search(x, n, b);
  double val = 0.0;
  for (int i = ia; i <= ib; ++i) {
    const double x0 = x[i];
    const double r1 = (i == ia) ? a - x0 : 0.0;
    const double r2 = (i == ib) ? b - x0 : x[i + 1] - x0;
    val += cubic_piece_integral(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3], r1, r2);
  }
  return val;
}

// log of the marginal integrand without the -A shift (pyx:84-138); -inf encodes
// "integrand is exactly zero"
__device__ __forceinline__ double log_integrand(const double* star, const double* data, int n,
                                                double SCALE, double B) {
  double x = 0.0;
  for (int j = 0; j < n; ++j) {
    const double c = SCALE * (star[j] + B);
    if (c > 0.0) x += data[j] * log(c) - c;
    else if (are_equal(c, 0.0) && are_equal(data[j], 0.0)) {}
    else return -INFINITY;
  }
  return x;
}

template <int BPL>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) k_marginal(MarginalArgs a) {
  constexpr int kMaxBins = 32 * BPL;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int chan = blockIdx.x * kWarpsPerBlock + warp;
  const int b = blockIdx.y;
  const int N_P = a.n_phases, n = a.n_bins;
  const int n_grids = a.comp_n_phases ? a.n_comp : 1;            // one phase grid per component, or one for all
  extern __shared__ double smem[];
  double* s_xall = smem;                                         // comp phases (shared by warps)
  double* w_base = smem + (long)n_grids * N_P + (long)warp * (6 * N_P + 2 * kMaxBins);
  double* s_y = w_base;                                          // [N_P]
  double* s_c = s_y + N_P;                                       // [N_P][4]
  double* s_m = s_c + 4 * N_P;                                   // [N_P] interval slopes (Akima)
  double* s_star = s_m + N_P;                                    // [kMaxBins]
  double* s_data = s_star + kMaxBins;                            // [kMaxBins]
  for (int i = threadIdx.x; i < n_grids * N_P; i += blockDim.x) s_xall[i] = a.comp_phases[i];
  __syncthreads();
  if (chan >= a.n_chan) return;            // whole warps only; no block barrier below

  const double nd = (double)n;
  const double T = a.exposure_time;
  const double SCALE = T / nd;
  const bool periodic = (a.interp != kSteffen);

  // ---- expected star count rate per bin (compute_expected_counts.pyx:66-197) ----------
  double star[BPL];
#pragma unroll
  for (int j = 0; j < BPL; ++j) star[j] = 0.0;
  for (int c = 0; c < a.n_comp; ++c) {
    const double* pulse = a.pulses + (((long)b * a.n_comp + c) * a.n_chan + chan) * N_P;
    const int Nc = a.comp_n_phases ? a.comp_n_phases[c] : N_P;   // nodes of this component's phase grid
    const double* s_x = s_xall + (a.comp_n_phases ? (long)c * N_P : 0);
    const bool allow_neg = a.comp_allow_negative ? (a.comp_allow_negative[c] != 0) : (a.allow_negative != 0);
    if (Nc == 1) {                     // time-invariant component: stored in the first bin, not integrated
      if (lane == 0) star[0] = pulse[0];                         // compute_expected_counts.pyx:190-192
      continue;
    }
    for (int i = lane; i < Nc; i += 32) s_y[i] = pulse[i];
    __syncwarp();
    if (a.interp == kCubic) {          // global C2 spline: one lane solves the cyclic system
      if (lane == 0) cspline_quads(s_x, s_y, Nc, true, s_c, 4);
    } else if (a.interp == kAkima) {
      // interval slopes once per interval (each is needed by five neighbours)
      for (int i = lane; i < Nc - 1; i += 32) s_m[i] = (s_y[i + 1] - s_y[i]) / (s_x[i + 1] - s_x[i]);
      __syncwarp();
      for (int i = lane; i < Nc - 1; i += 32) {
        double bb, cc, dd;
        akima_coeffs_cached(s_m, s_x, s_y, Nc, i, periodic, &bb, &cc, &dd);
        s_c[4 * i] = s_y[i]; s_c[4 * i + 1] = bb; s_c[4 * i + 2] = cc; s_c[4 * i + 3] = dd;
      }
    } else
    for (int i = lane; i < Nc - 1; i += 32) {
      double bb, cc, dd;
      interp_coeffs(a.interp, periodic, s_x, s_y, Nc, i, &bb, &cc, &dd);
      s_c[4 * i] = s_y[i]; s_c[4 * i + 1] = bb; s_c[4 * i + 2] = cc; s_c[4 * i + 3] = dd;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < BPL; ++j) {
      const int bin = lane + 32 * j;
      if (bin >= n) continue;
      const double shift = a.phase_shifts[(long)b * a.n_comp + c];
      double pa = a.data_phases[bin] + shift;
      double pb = a.data_phases[bin + 1] + shift;
      if (are_equal(pb - pa, 1.0)) { pa = 0.0; pb = 1.0; }
      else { pa -= floor(pa); pb -= floor(pb); }
      if (pa < pb) {
        const double v = spline_integ(s_x, s_c, Nc, pa, pb);
        if (v > 0.0 || allow_neg) star[j] += v;
      } else {
        double v = spline_integ(s_x, s_c, Nc, pa, 1.0);
        if (v > 0.0 || allow_neg) star[j] += v;
        v = spline_integ(s_x, s_c, Nc, 0.0, pb);
        if (v > 0.0 || allow_neg) star[j] += v;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < BPL; ++j) if (star[j] < 0.0) star[j] = 0.0;
  if (a.given_background) {
    // compute_expected_counts.pyx:300-305 + _poisson_likelihood_given_background.pyx:96-113
    double term = 0.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < BPL; ++j) {
      const int bin = lane + 32 * j;
      if (bin >= n) continue;
      const double expec = (star[j] + a.background[(long)chan * n + bin]) * T;
      const double cnt = a.counts ? a.counts[(long)chan * n + bin] : 0.0;
      if (expec > 0.0) term += cnt * log(expec) - expec;
      else if (cnt == 0.0 && expec == 0.0) {}
      else bad = 1;
      if (a.expected) a.expected[((long)b * a.n_chan + chan) * n + bin] = expec;
    }
    const double sum = warp_sum(term);
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) {
      a.chan_lnL[(long)b * a.n_chan + chan] = sum + (a.precomp ? a.precomp[chan] : 0.0);
      a.chan_status[(long)b * a.n_chan + chan] = anybad ? 2 : 0;
    }
    return;
  }
  double d[BPL];
  double sum_star = 0.0, sum_d = 0.0;
#pragma unroll
  for (int j = 0; j < BPL; ++j) {
    const int bin = lane + 32 * j;
    d[j] = 0.0;
    if (bin < n) {
      if (a.background) star[j] += a.background[(long)chan * n + bin];
      star[j] *= nd;                                           // pyx:659-665
      d[j] = a.counts[(long)chan * n + bin];
    } else star[j] = 0.0;
    s_star[bin] = star[j]; s_data[bin] = d[j];
    sum_star += star[j]; sum_d += d[j];
  }
  __syncwarp();
  double av_STAR = warp_sum(sum_star);
  double av_DATA = warp_sum(sum_d);
  // sums over this lane's bins of f(star, data)
  auto bins_sum = [&](auto f) -> double {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < BPL; ++j) if (lane + 32 * j < n) acc += f(star[j], d[j]);
    return acc;
  };

  int status = 0;
  double loglike = 0.0, B = 0.0;
  const double sup0 = a.support[2 * chan], sup1 = a.support[2 * chan + 1];

  if (a.slim >= 0.0) {                                         // pyx:677-684
    const double limit = av_STAR * SCALE - a.slim * sqrt(av_STAR * SCALE) - av_DATA;
    if (limit > 0.0) status = 1;
  }
  av_STAR /= nd;
  av_DATA /= T;

  if (status == 0) {
    if (are_equal(av_DATA, 0.0) && are_equal(av_STAR, 0.0)) {  // pyx:287-301
      double lower = 0.0;
      if (lower < sup0) lower = sup0;
      double upper = 10.0 / T;
      if (upper > sup1 && sup1 > 0.0) upper = sup1;
      B = 0.0;
      loglike = log((exp(-1.0 * lower * T) - exp(-1.0 * upper * T)) / T);
    } else {
      double B_min = 0.0;
      B = av_DATA - av_STAR;
      if (B <= B_min) {                                        // pyx:309-332
        bool any_zero = false;
        double m = INFINITY;                      // smallest positive count among the bins
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
          if (lane + 32 * j >= n) continue;
          if (are_equal(star[j], 0.0)) any_zero = true;
          if (d[j] > 0.0) m = fmin(m, d[j]);
        }
        const unsigned zero_star = __ballot_sync(0xffffffffu, any_zero);
        if (zero_star) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (isinf(m)) { B = 0.0; B_min = 0.0; }
          else { B = 0.01 * m / SCALE; B_min = 0.1 * B; }
        } else {
          B = B_min;
        }
      }
      // Newton iterations (pyx:336-352); delta() = pyx:140-173
      double std, dB;
      {
        const double y = warp_sum(bins_sum([&](double s_, double d_) { return d_ / (s_ + B); }));
        const double x = warp_sum(bins_sum([&](double s_, double d_) { return d_ / (s_ + B) / (s_ + B); }));
        std = sqrt(2.0 / (2.0 * x));
        dB = -1.0 * (2.0 * T - 2.0 * y) / (2.0 * x);
      }
      int counter = 0, iters = 0;
      while (fabs(dB) > a.epsilon * std && counter < 2) {
        B += dB;
        if (B < B_min) {
          if (B_min > 0.0) counter += 1; else counter = 2;
          B = B_min;
        }
        const double y = warp_sum(bins_sum([&](double s_, double d_) { return d_ / (s_ + B); }));
        const double x = warp_sum(bins_sum([&](double s_, double d_) { return d_ / (s_ + B) / (s_ + B); }));
        std = sqrt(2.0 / (2.0 * x));
        dB = -1.0 * (2.0 * T - 2.0 * y) / (2.0 * x);
        if (++iters > 500) { status = 2; break; }
      }
      double std_est = warp_sum(bins_sum([&](double s_, double d_) { return d_ / ((s_ + B) * (s_ + B)); }));
      std_est = std_est > 0.0 ? sqrt(1.0 / std_est) : 1e90;
      double lower = B - a.sigmas * std_est;
      double upper = B + a.sigmas * std_est;
      if (lower < B_min) lower = B_min;
      double Bfi = B;                                          // pyx:366-393
      if (lower < sup0) {
        lower = sup0;
        if (upper < sup0 && sup1 > 0.0) { upper = sup1; Bfi = sup0; }
        else if (upper < sup0) { upper = sup0 + a.sigmas * std_est; Bfi = sup0; }
      }
      if (upper > sup1 && sup1 > 0.0) {
        upper = sup1;
        if (lower > sup1) { lower = sup0; Bfi = sup1; }
      }
      if (Bfi < lower) Bfi = lower; else if (Bfi > upper) Bfi = upper;
      // A: log-integrand at the clipped ML background (pyx:397-408)
      double A;
      {
        const double llz = a.llzero;
        A = warp_sum(bins_sum([&](double s_, double d_) {
          const double c = SCALE * (s_ + Bfi);
          if (c > 0.0) return d_ * log(c) - c;
          if (are_equal(c, 0.0) && are_equal(d_, 0.0)) return 0.0;
          return llz;
        }));
      }
      // ---- marginal integral over [lower, upper] (replaces pyx:416-419) -------------
      double result = 0.0;
      if (status == 0 && upper > lower) {
        double lo = lower, hi = upper;
        for (int zoom = 0; zoom < 6; ++zoom) {
          const double h = (hi - lo) / 31.0;
          const double Bp = (lane == 31) ? hi : lo + h * lane;
          const double xl = log_integrand(s_star, s_data, n, SCALE, Bp);
          const double ref = fmax(warp_max(xl), A);
          const unsigned mask = __ballot_sync(0xffffffffu, xl >= ref - 45.0);
          if (mask == 0u) {              // peak narrower than the probe spacing: centre on Bfi
            lo = fmax(lo, Bfi - h); hi = fmin(hi, Bfi + h);
            continue;
          }
          const int first = __ffs(mask) - 1, last = 31 - __clz(mask);
          const double nlo = (first > 0) ? lo + h * (first - 1) : lo;
          const double nhi = (last < 31) ? ((last + 1 == 31) ? hi : lo + h * (last + 1)) : hi;
          const bool resolved = (last - first) >= 8;
          lo = nlo; hi = nhi;
          if (resolved) break;
        }
        // kPanels (12) panels x 8-point Gauss-Legendre over the bracketed set: the integrand is close to a Gaussian
        // about 19 sigma wide there, so a panel spans ~1.6 sigma and the 8-point rule's error term is ~3e-14 of the
        // integral (16 panels: 1e-17, 8 panels: 3e-11; measured on 2 300 evaluations in deterministic-free runs: 12 vs
        // 16 panels differ by <= 1.2e-9 in lnL at |lnL| ~ 1e5-1e6, i.e. the level of the ring-sum jitter)
        const double hp = (hi - lo) / (double)kPanels;
        double acc = 0.0;
#pragma unroll 1
        for (int batch = 0; batch < kPanels / 4; ++batch) {
          const int node = batch * 32 + lane;
          const int panel = node >> 3, g = node & 7;
          const double Bq = lo + hp * ((double)panel + 0.5 + 0.5 * c_gl8_x[g]);
          const double xq = log_integrand(s_star, s_data, n, SCALE, Bq);
          acc += 0.5 * hp * c_gl8_w[g] * exp(xq - A);           // exp(-inf) = 0
        }
        result = warp_sum(acc);
      }
      if (result > 0.0) loglike = log(result) + A + a.precomp[chan];   // pyx:423-424
      else if (status == 0) status = 2;
    }
  }

  // ---- outputs (pyx:431-447) --------------------------------------------------------
  if (lane == 0) {
    a.chan_lnL[(long)b * a.n_chan + chan] = loglike;
    a.chan_status[(long)b * a.n_chan + chan] = status;
    if (a.mcl_bg) a.mcl_bg[(long)b * a.n_chan + chan] = B * T;
  }
  double Bc = B;
  if (Bc < sup0) Bc = sup0; else if (Bc > sup1 && sup1 > 0.0) Bc = sup1;
  if (lane == 0 && a.mcl_bg_support) a.mcl_bg_support[(long)b * a.n_chan + chan] = Bc * T;
  if (a.expected) {
#pragma unroll
    for (int j = 0; j < BPL; ++j)
      if (lane + 32 * j < n)
        a.expected[((long)b * a.n_chan + chan) * n + lane + 32 * j] = (status == 0) ? SCALE * (star[j] + Bc) : star[j];
  }
}

// ordered reduction over channels; the first failing channel decides the status
// (the reference breaks out of its channel loop there, pyx:681-684,716-719)
__global__ void k_sum_channels(const double* chan_lnL, const int* chan_status, int n_chan,
                               double* lnL, int* status) {
  const int b = blockIdx.x;
  __shared__ double s_sum[256];
  __shared__ int s_bad[256];
  double s = 0.0;
  int bad = n_chan;
  for (int c = threadIdx.x; c < n_chan; c += blockDim.x) {
    s += chan_lnL[(long)b * n_chan + c];
    if (chan_status[(long)b * n_chan + c] != 0 && c < bad) bad = c;
  }
  s_sum[threadIdx.x] = s; s_bad[threadIdx.x] = bad;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_bad[threadIdx.x] = min(s_bad[threadIdx.x], s_bad[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    lnL[b] = s_sum[0];
    // a status already raised upstream (integrator failure) is kept
    if (status[b] == 0) status[b] = (s_bad[0] < n_chan) ? 10 + chan_status[(long)b * n_chan + s_bad[0]] : 0;
  }
}

template <int BPL>
static cudaError_t launch_marginal_bpl(const MarginalArgs& a, cudaStream_t stream) {
  const size_t n_grids = a.comp_n_phases ? a.n_comp : 1;
  const size_t smem = (n_grids * a.n_phases + kWarpsPerBlock * (6ul * a.n_phases + 2 * 32 * BPL)) * sizeof(double);
  dim3 grid((a.n_chan + kWarpsPerBlock - 1) / kWarpsPerBlock, a.B);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_marginal<BPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_marginal<BPL><<<grid, 32 * kWarpsPerBlock, smem, stream>>>(a);
  return cudaGetLastError();
}

int marginal_max_bins() { return 32 * kMaxBPL; }

cudaError_t launch_marginal(MarginalArgs a, cudaStream_t stream) {
  if (a.n_bins > 32 * kMaxBPL || a.n_bins < 1) return cudaErrorNotSupported;
  if (a.n_phases < 5 && a.n_phases != 1) return cudaErrorInvalidValue;     // 1: time-invariant components
  if (a.interp == kCubic && a.n_phases > kMaxCubicNodes) return cudaErrorInvalidValue;
  cudaError_t err = a.n_bins <= 32 ? launch_marginal_bpl<1>(a, stream)
                    : a.n_bins <= 64 ? launch_marginal_bpl<2>(a, stream) : launch_marginal_bpl<4>(a, stream);
  if (err != cudaSuccess) return err;
  if (a.lnL) {
    k_sum_channels<<<a.B, 256, 0, stream>>>(a.chan_lnL, a.chan_status, a.n_chan, a.lnL, a.status);
    err = cudaGetLastError();
  }
  return err;
}

}  // namespace xb
