// obca_traj.h -- trajectory-side kernels around the solve:
//
//   [INTERP]   Vehicle.get_interpolator / interpolate_states (confrez/control/vehicle.py:722-829): the collocation solution
//              evaluated at arbitrary times -- states by the degree-K Lagrange polynomial of the interval that contains t,
//              inputs piecewise constant between the collocation nodes, the final state held beyond the horizon.
//   [REF]      VehicleFollower.get_current_ref (confrez/control/vehicle_follower.py:370-404): nearest sample of the dense
//              reference time grid to the vehicle's clock, then N samples dt apart from there.
//   [PLANT]    simulator (confrez/control/dynamic_model.py:61-93): the bicycle ODE over one control period under constant
//              inputs.  The reference integrates with IDAS; here fixed-step RK4 (100 sub-steps: local error < 1e-12).
//   [SHIFT]    VehicleFollower._adv_onestep (vehicle_follower.py:413-426): one-step shift with last-value hold.
//
// One thread per output sample; every function is written once for the kernels and for the host emulation.
#pragma once

namespace obca {

// state / input sample of vehicle (b, a) at time t.  z: (B,V,Mmax,7) node major, dtv: interval length of instance b,
// N: intervals of this vehicle, tau[6]: collocation nodes on [0, 1].  out[7] = x y psi v delta a w
OBCA_HD void traj_sample(const double* z, double dtv, int N, const double* tau, double t, double* out) {
  // interval: tgrid = linspace(0, N dt, N + 1); pw_const picks interval i while t < tgrid[i + 1]
  const double step = (N * dtv) / N;
  auto tg = [&](int k) { return k < N ? k * step : N * dtv; };  // numpy.linspace: start + k * step, the last point is `stop` itself
  int i = (int)floor(t / step);
  if (i < 0) i = 0;
  if (i > N) i = N;
  while (i < N && tg(i + 1) <= t) ++i;
  while (i > 0 && tg(i) > t) --i;
  if (i >= N) {
    // beyond the horizon every collocation value is the final state lf = sum_k D_k X[N-1][k] = the last node (D = e_K)
    const double* zl = z + (size_t)(N * NK - 1) * NZ;
    for (int q = 0; q < 5; ++q) out[q] = zl[q];
  } else {
    const double rel = (t - tg(i)) / dtv;
    double basis[NK];
    for (int j = 0; j < NK; ++j) {
      double b = 1.0;
      for (int k = 0; k < NK; ++k)
        if (k != j) b *= (rel - tau[k]) / (tau[j] - tau[k]);
      basis[j] = b;
    }
    for (int q = 0; q < 5; ++q) {
      double s = 0;
      for (int j = 0; j < NK; ++j) s += basis[j] * z[(size_t)(i * NK + j) * NZ + q];
      out[q] = s;
    }
  }
  // inputs: pw_const(t, t_nodes[1:], u): index = #(t_nodes[1:] <= t), clipped; t_nodes[m] = (m / NK + tau[m % NK]) dt
  const int M = N * NK;
  auto tn = [&](int m) { return (m / NK + tau[m % NK]) * dtv; };
  int m = i < N ? i * NK : M - 1;
  while (m > 0 && tn(m) > t) --m;
  while (m < M - 1 && tn(m + 1) <= t) ++m;
  out[5] = z[(size_t)m * NZ + 5];
  out[6] = z[(size_t)m * NZ + 6];
}

struct InterpArgs {
  const double *z, *dt, *times, *tau;  // times: (T) shared by all instances, or (B,V,T) when per_vehicle_times
  const int* n_intervals;              // (V)
  double* out;                         // (B,V,T,7)
  int B, V, Mmax, T, per_vehicle_times, dt_per_vehicle;  // dt: (B), or (B,V) when dt_per_vehicle (every vehicle has its own plan)
};

OBCA_HD void interp_item(const InterpArgs& A, size_t g) {
  const int t = (int)(g % A.T), a = (int)((g / A.T) % A.V);
  const size_t b = g / ((size_t)A.T * A.V);
  const double tq = A.per_vehicle_times ? A.times[g] : A.times[t];
  const double dtv = A.dt_per_vehicle ? A.dt[b * A.V + a] : A.dt[b];
  traj_sample(A.z + ((b * A.V + a) * (size_t)A.Mmax) * NZ, dtv, A.n_intervals[a], A.tau, tq, A.out + g * NZ);
}

// [REF] times of the MPC reference window of vehicle (b): t_ref[argmin |t_ref - clock|] + k dt, k < N; the dense grid is
// linspace(t_first, t_last, n_ref) (vehicle_follower.py:130-137), ties resolve to the lower index like numpy.argmin
OBCA_HD double ref_window_start(double t_first, double t_last, int n_ref, double clock) {
  if (n_ref <= 1) return t_first;
  const double step = (t_last - t_first) / (n_ref - 1);
  int k = (int)floor((clock - t_first) / step);
  if (k < 0) k = 0;
  if (k > n_ref - 1) k = n_ref - 1;
  int best = k;
  double bd = fabs(t_first + k * step - clock);
  for (int c = k - 1; c <= k + 1; ++c) {
    if (c < 0 || c > n_ref - 1) continue;
    const double d = fabs(t_first + c * step - clock);
    if (d < bd || (d == bd && c < best)) bd = d, best = c;
  }
  return t_first + best * step;
}

// [PLANT] RK4 with `substeps` sub-steps over dt under constant input
OBCA_HD void plant_step(const double* s, const double* u, double dt, double wb, int substeps, double* out) {
  double z[5] = {s[0], s[1], s[2], s[3], s[4]};
  const double h = dt / substeps;
  for (int it = 0; it < substeps; ++it) {
    double k1[5], k2[5], k3[5], k4[5], w[5];
    auto f = [&](const double* q, double* d) {
      d[0] = q[3] * cos(q[2]), d[1] = q[3] * sin(q[2]), d[2] = q[3] / wb * tan(q[4]), d[3] = u[0], d[4] = u[1];
    };
    f(z, k1);
    for (int q = 0; q < 5; ++q) w[q] = z[q] + h / 2 * k1[q];
    f(w, k2);
    for (int q = 0; q < 5; ++q) w[q] = z[q] + h / 2 * k2[q];
    f(w, k3);
    for (int q = 0; q < 5; ++q) w[q] = z[q] + h * k3[q];
    f(w, k4);
    for (int q = 0; q < 5; ++q) z[q] = z[q] + h / 6 * (k1[q] + 2 * k2[q] + 2 * k3[q] + k4[q]);
  }
  for (int q = 0; q < 5; ++q) out[q] = z[q];
}

// [WSINTERP] Vehicle.interp_ws_for_collocation (confrez/control/vehicle.py:298-358): every warm-start signal, sampled on the grid
// t (T), is interpolated linearly (scipy interp1d) onto the collocation times t_interp[i (K+1) + k] = (i + tau_k) / N * t[T-1].
// in (B,T,C) -> out (B,N (K+1),C); one thread per output entry.
OBCA_HD void ws_interp_item(const double* in, const double* t, const double* tau, int T, int C, int N, double* out, size_t g) {
  const int c = (int)(g % C), m = (int)((g / C) % (N * NK));
  const size_t b = g / ((size_t)C * N * NK);
  const double tq = ((m / NK) + tau[m % NK]) / N * t[T - 1];
  int lo = 0, hi = T - 1;  // largest lo with t[lo] <= tq, kept below T - 1 (tq never exceeds t[T-1]: tau <= 1, i <= N - 1)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t[mid] <= tq) lo = mid;
    else hi = mid;
  }
  const double* s = in + (b * T) * (size_t)C + c;
  const double y0 = s[(size_t)lo * C], y1 = s[(size_t)(lo + 1) * C];
  // scipy.interpolate.interp1d(kind="linear"): slope * (x_new - x_lo) + y_lo
  out[g] = (y1 - y0) / (t[lo + 1] - t[lo]) * (tq - t[lo]) + y0;
}

// [SHIFT] out[b][n][:] = in[b][min(n + 1, N - 1)][:] for (B, N, W) arrays
OBCA_HD void shift_item(const double* in, double* out, int N, int Wd, size_t g) {
  const int w = (int)(g % Wd), n = (int)((g / Wd) % N);
  const size_t b = g / ((size_t)Wd * N);
  const int src = n + 1 < N ? n + 1 : N - 1;
  out[g] = in[(b * N + src) * (size_t)Wd + w];
}

}  // namespace obca
